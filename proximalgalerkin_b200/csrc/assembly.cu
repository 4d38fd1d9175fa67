// Element kernels, atomic-free row gather, the 2x2-block sliced-ELL operator (J*v and residual),
// observables and CSR export.
//
// What the reference evaluates through FFCx-generated tabulate_tensor kernels + dolfinx assemblers
// for the forms of examples/01_obstacle_problem/obstacle_pg.py:116-125, restructured around the
// block form  J = [[alpha K, M], [M, -D(psi)]]  (explicit in obstacle_finite_difference.jl:37-38):
// K, M and the obstacle/forcing load vectors are assembled once; per Newton step only
// D(psi) = int exp(psi) phi_i phi_j is re-assembled, and  int exp(psi) phi_i = (D 1)_i  by partition
// of unity, so the residual is one pass over the stored operator.
#include "lvpp_internal.cuh"
#include "block_op.cuh"

// closed form of obstacle_pg.py:92-104
__device__ __forceinline__ double phi_set(double r) {
  const double r0 = 0.5, beta = 0.9;
  const double b = r0 * beta;
  const double tmp = sqrt(r0 * r0 - b * b);
  const double B = tmp + b * b / tmp;
  const double Cc = -b / tmp;
  return r > b ? B + r * Cc : sqrt(fmax(r0 * r0 - r * r, 0.0));
}

// Row a of the symmetric element matrix (packed upper triangle `acc`, times `scale`) goes to the
// incidence-ELL slot of (cell, a): NLD consecutive doubles, so the row gather streams them coalesced.
template <int NLD>
__device__ __forceinline__ void store_element_rows(const double* acc, double scale, const uint32_t* __restrict__ pos,
                                                   double* __restrict__ De) {
#pragma unroll
  for (int a = 0; a < NLD; ++a) {
    const uint32_t p = pos[a];
    if (p == 0xffffffffu) continue;  // row of a ghost node: assembled by its owner
    double* dst = De + (int64_t)p * NLD;
    if (NLD % 2 == 0) {
#pragma unroll
      for (int b = 0; b < NLD; b += 2) {
        const double v0 = scale * acc[a <= b ? lvpp_sym(a, b, NLD) : lvpp_sym(b, a, NLD)];
        const double v1 = scale * acc[a <= b + 1 ? lvpp_sym(a, b + 1, NLD) : lvpp_sym(b + 1, a, NLD)];
        *reinterpret_cast<double2*>(dst + b) = make_double2(v0, v1);
      }
    } else {
#pragma unroll
      for (int b = 0; b < NLD; ++b) dst[b] = scale * acc[a <= b ? lvpp_sym(a, b, NLD) : lvpp_sym(b, a, NLD)];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Setup cell kernel.  which = 0: stiffness K_e, 1: mass M_e into De (packed upper triangle);
// which = 2: |det J| and the load vectors  int phi_obs phi_a  and  int phi_a  into ve[C][2][NLD].
template <int TDIM, int NLD>
__global__ void __launch_bounds__(128)
k_cell_setup(int which, int64_t C, int nq, const int32_t* __restrict__ cells,
             const double* __restrict__ coords, const double* __restrict__ tab,
             int obstacle_kind, double period, double origin, const double* __restrict__ phi_obs_q,
             const uint32_t* __restrict__ pos,
             double* __restrict__ De, double* __restrict__ adetJ, double* __restrict__ ve) {
  constexpr int NSYM = NLD * (NLD + 1) / 2;
  extern __shared__ double s_tab[];
  stage_tables(s_tab, tab, nq * (1 + NLD + NLD * TDIM + TDIM));
  const double* s_w = s_tab;
  const double* s_phi = s_w + nq;
  const double* s_dphi = s_phi + nq * NLD;
  const double* s_qp = s_dphi + nq * NLD * TDIM;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < C;
       c += (int64_t)gridDim.x * blockDim.x) {
    int32_t nodes[NLD];
#pragma unroll
    for (int a = 0; a < NLD; ++a) nodes[a] = cells[c * NLD + a];
    double Jinv[TDIM][TDIM];
    const double adet = cell_geometry<TDIM>(coords, nodes, Jinv);
    if (which == 2) {
      double xv[TDIM + 1][TDIM];
#pragma unroll
      for (int v = 0; v <= TDIM; ++v)
#pragma unroll
        for (int d = 0; d < TDIM; ++d) xv[v][d] = coords[(int64_t)nodes[v] * TDIM + d];
      double lo[NLD], lf[NLD];
#pragma unroll
      for (int a = 0; a < NLD; ++a) lo[a] = lf[a] = 0.0;
      for (int q = 0; q < nq; ++q) {
        double po;
        if (obstacle_kind == LVPP_OBSTACLE_PHI_SET) {
          double lam0 = 1.0, r2 = 0.0;
#pragma unroll
          for (int d = 0; d < TDIM; ++d) lam0 -= s_qp[q * TDIM + d];
#pragma unroll
          for (int d = 0; d < TDIM; ++d) {
            double xq = lam0 * xv[0][d];
#pragma unroll
            for (int v = 1; v <= TDIM; ++v) xq += s_qp[q * TDIM + v - 1] * xv[v][d];
            if (d == TDIM - 1 && period > 0.0) {  // one obstacle per period of the stacked domain
              xq = fmod(xq - origin, period);
              if (xq < 0.0) xq += period;
              xq -= 0.5 * period;
            }
            r2 += xq * xq;
          }
          po = phi_set(sqrt(r2));
        } else {
          po = phi_obs_q[c * nq + q];
        }
        const double w = s_w[q] * adet;
#pragma unroll
        for (int a = 0; a < NLD; ++a) {
          const double p = s_phi[q * NLD + a];
          lo[a] += w * po * p;
          lf[a] += w * p;
        }
      }
      adetJ[c] = adet;
#pragma unroll
      for (int a = 0; a < NLD; ++a) {
        ve[(c * 2 + 0) * NLD + a] = lo[a];
        ve[(c * 2 + 1) * NLD + a] = lf[a];
      }
      continue;
    }
    double acc[NSYM];
#pragma unroll
    for (int s = 0; s < NSYM; ++s) acc[s] = 0.0;
    for (int q = 0; q < nq; ++q) {
      const double w = s_w[q] * adet;
      if (which == 0) {
        double g[NLD][TDIM];  // physical gradients: g_a = J^{-T} dphi_a
#pragma unroll
        for (int a = 0; a < NLD; ++a)
#pragma unroll
          for (int d = 0; d < TDIM; ++d) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < TDIM; ++k) s += s_dphi[(q * NLD + a) * TDIM + k] * Jinv[k][d];
            g[a][d] = s;
          }
        int s = 0;
#pragma unroll
        for (int a = 0; a < NLD; ++a)
#pragma unroll
          for (int b = a; b < NLD; ++b, ++s) {
            double dot = 0.0;
#pragma unroll
            for (int d = 0; d < TDIM; ++d) dot += g[a][d] * g[b][d];
            acc[s] += w * dot;
          }
      } else {
        int s = 0;
#pragma unroll
        for (int a = 0; a < NLD; ++a)
#pragma unroll
          for (int b = a; b < NLD; ++b, ++s) acc[s] += w * s_phi[q * NLD + a] * s_phi[q * NLD + b];
      }
    }
    store_element_rows<NLD>(acc, 1.0, pos + c * NLD, De);
  }
}

// Per-Newton-step cell kernel: D_e = |det J| sum_q w_q exp(psi(x_q)) phi_a phi_b (packed upper triangle)
template <int NLD>
__global__ void __launch_bounds__(128)
k_cell_exp(int64_t C, int nq, const int32_t* __restrict__ cells, const double2* __restrict__ x,
           const double* __restrict__ adetJ, const double* __restrict__ tab, const uint32_t* __restrict__ pos,
           double* __restrict__ De) {
  constexpr int NSYM = NLD * (NLD + 1) / 2;
  extern __shared__ double s_tab[];
  stage_tables(s_tab, tab, nq * (1 + NLD));
  const double* s_w = s_tab;
  const double* s_phi = s_w + nq;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < C;
       c += (int64_t)gridDim.x * blockDim.x) {
    double psi[NLD];
#pragma unroll
    for (int a = 0; a < NLD; ++a) psi[a] = __ldg(&x[cells[c * NLD + a]]).y;
    double acc[NSYM];
#pragma unroll
    for (int s = 0; s < NSYM; ++s) acc[s] = 0.0;
    for (int q = 0; q < nq; ++q) {
      double pq = 0.0;
#pragma unroll
      for (int a = 0; a < NLD; ++a) pq += s_phi[q * NLD + a] * psi[a];
      const double e = exp(pq) * s_w[q];
      int s = 0;
#pragma unroll
      for (int a = 0; a < NLD; ++a) {
        const double ea = e * s_phi[q * NLD + a];
#pragma unroll
        for (int b = a; b < NLD; ++b, ++s) acc[s] += ea * s_phi[q * NLD + b];
      }
    }
    store_element_rows<NLD>(acc, adetJ[c], pos + c * NLD, De);
  }
}

// Row gather (the atomic-free scatter): one thread per owned node sums the element rows of its
// incident cells, in cell order, into per-thread shared-memory accumulators and writes its SELL row.
// The element rows arrive in incidence-ELL order: at step t the 32 lanes of a warp read 32 consecutive
// rows (NLD doubles each) and their slot-offset bytes -- fully coalesced streams.  Deterministic: the
// summation order is fixed by the incidence list.
template <int NLD>
__global__ void k_row_gather(int64_t Vown, const int64_t* __restrict__ inc_ptr, const int64_t* __restrict__ ie_ptr,
                             const uint8_t* __restrict__ ie_k, const double* __restrict__ De,
                             const int64_t* __restrict__ slice_ptr, const int32_t* __restrict__ rowlen,
                             double* __restrict__ out) {
  constexpr int KB = (NLD + 3) & ~3;
  extern __shared__ double s_acc[];  // [maxw][blockDim]
  const int nt = blockDim.x, tid = threadIdx.x;
  for (int64_t i0 = blockIdx.x * (int64_t)nt; i0 < Vown; i0 += (int64_t)gridDim.x * nt) {
    const int64_t i = i0 + tid;
    if (i >= Vown) continue;
    const int len = rowlen[i];
    for (int k = 0; k < len; ++k) s_acc[k * nt + tid] = 0.0;
    const int cnt = (int)(inc_ptr[i + 1] - inc_ptr[i]);
    const int64_t ibase = ie_ptr[i >> 5] + (i & 31);
    for (int t = 0; t < cnt; ++t) {
      const int64_t slot = ibase + (int64_t)t * LVPP_SLICE;
      const double* de = De + slot * NLD;
      uint8_t kk[KB];
#pragma unroll
      for (int q = 0; q < KB / 4; ++q) {
        const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(ie_k + slot * KB) + q);
        kk[4 * q] = u.x; kk[4 * q + 1] = u.y; kk[4 * q + 2] = u.z; kk[4 * q + 3] = u.w;
      }
      double v[NLD];
      if (NLD % 2 == 0) {
#pragma unroll
        for (int b = 0; b < NLD; b += 2) {
          const double2 d2 = __ldcs(reinterpret_cast<const double2*>(de + b));
          v[b] = d2.x; v[b + 1] = d2.y;
        }
      } else {
#pragma unroll
        for (int b = 0; b < NLD; ++b) v[b] = __ldcs(de + b);
      }
#pragma unroll
      for (int b = 0; b < NLD; ++b) s_acc[kk[b] * nt + tid] += v[b];
    }
    const int64_t base = slice_ptr[i >> 5] + (i & 31);
    for (int k = 0; k < len; ++k) out[base + (int64_t)k * LVPP_SLICE] = s_acc[k * nt + tid];
  }
}

// node-vector gather of ve[C][2][NLD] (obstacle and unit load vectors)
template <int NLD>
__global__ void k_vec_gather(int64_t Vown, const int64_t* __restrict__ inc_ptr,
                             const uint32_t* __restrict__ inc_val, const double* __restrict__ ve,
                             double* __restrict__ bobs, double* __restrict__ fvec) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown;
       i += (int64_t)gridDim.x * blockDim.x) {
    double s0 = 0.0, s1 = 0.0;
    const int64_t e1 = inc_ptr[i + 1];
    for (int64_t e = inc_ptr[i]; e < e1; ++e) {
      const uint32_t v = inc_val[e];
      const int64_t c = v / NLD;
      const int a = (int)(v - (uint32_t)c * NLD);
      s0 += ve[(c * 2 + 0) * NLD + a];
      s1 += ve[(c * 2 + 1) * NLD + a];
    }
    bobs[i] = s0;
    fvec[i] = s1;
  }
}

// deterministic second stage: sums nvals interleaved partial arrays [nvals][nparts]
__global__ void __launch_bounds__(256) k_reduce_partials(int nparts, int nvals, const double* __restrict__ partials,
                                                          double* __restrict__ out) {
  __shared__ double s_red[32];
  for (int v = 0; v < nvals; ++v) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 256) s += partials[(int64_t)v * nparts + i];
    const double r = lvpp_block_sum<256>(s, s_red);
    if (threadIdx.x == 0) out[v] = r;
  }
}

// ------------------------------------------------------------------------------------------------
// Observables of obstacle_pg.py:145-152 (cell loop with the form quadrature); partial sums per block.
template <int TDIM, int NLD>
__global__ void __launch_bounds__(128)
k_observables(int64_t Cown, int nq, const int32_t* __restrict__ cells, const double* __restrict__ coords,
              const double* __restrict__ tab, const double2* __restrict__ x, const double2* __restrict__ xk,
              double alpha, double f, int nparts, double* __restrict__ partials) {
  extern __shared__ double s_tab[];
  __shared__ double s_red[32];
  stage_tables(s_tab, tab, nq * (1 + NLD + NLD * TDIM));
  const double* s_w = s_tab;
  const double* s_phi = s_w + nq;
  const double* s_dphi = s_phi + nq * NLD;
  double o[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < Cown;
       c += (int64_t)gridDim.x * blockDim.x) {
    int32_t nodes[NLD];
    double u[NLD], ps[NLD], uk[NLD], pk[NLD];
#pragma unroll
    for (int a = 0; a < NLD; ++a) {
      nodes[a] = cells[c * NLD + a];
      const double2 v = __ldg(&x[nodes[a]]), vk = __ldg(&xk[nodes[a]]);
      u[a] = v.x; ps[a] = v.y; uk[a] = vk.x; pk[a] = vk.y;
    }
    double Jinv[TDIM][TDIM];
    const double adet = cell_geometry<TDIM>(coords, nodes, Jinv);
    for (int q = 0; q < nq; ++q) {
      double uq = 0, pq = 0, ukq = 0, pkq = 0, gu[TDIM], gd[TDIM];
#pragma unroll
      for (int d = 0; d < TDIM; ++d) gu[d] = gd[d] = 0.0;
#pragma unroll
      for (int a = 0; a < NLD; ++a) {
        const double ph = s_phi[q * NLD + a];
        uq += ph * u[a]; pq += ph * ps[a]; ukq += ph * uk[a]; pkq += ph * pk[a];
#pragma unroll
        for (int d = 0; d < TDIM; ++d) {
          double g = 0.0;
#pragma unroll
          for (int k = 0; k < TDIM; ++k) g += s_dphi[(q * NLD + a) * TDIM + k] * Jinv[k][d];
          gu[d] += g * u[a];
          gd[d] += g * (u[a] - uk[a]);
        }
      }
      const double w = s_w[q] * adet;
      double gu2 = 0, gd2 = 0;
#pragma unroll
      for (int d = 0; d < TDIM; ++d) { gu2 += gu[d] * gu[d]; gd2 += gd[d] * gd[d]; }
      o[0] += w * (0.5 * gu2 - f * uq);
      o[1] += w * ((pkq - pq) / alpha * uq);
      o[2] += w * (uq < 0.0 ? -uq : 0.0);
      o[3] += w * (pkq < pq ? (pq - pkq) / alpha : 0.0);
      o[4] += w * (gd2 + (uq - ukq) * (uq - ukq));
      const double de = exp(pq) - exp(pkq);
      o[5] += w * de * de;
    }
  }
#pragma unroll
  for (int v = 0; v < 6; ++v) {
    const double r = lvpp_block_sum<128>(o[v], s_red);
    if (threadIdx.x == 0) partials[(int64_t)v * nparts + blockIdx.x] = r;
  }
}

// values of the assembled Jacobian on the monolithic CSR pattern, Dirichlet rows/columns applied
__global__ void k_export_values(int64_t Vown, const int64_t* __restrict__ rowptr,
                                const int64_t* __restrict__ slice_ptr, const int32_t* __restrict__ rowlen,
                                const uint32_t* __restrict__ col, const double* __restrict__ K,
                                const double* __restrict__ M, const double* __restrict__ D,
                                const uint8_t* __restrict__ bc_flag, double alpha, double* __restrict__ vals) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t S = rowptr[i];
    const int len = rowlen[i];
    const int64_t p0 = 4 * S, p1 = 4 * S + 2 * len;
    const int64_t base = slice_ptr[i >> 5] + (i & 31);
    const bool rbc = bc_flag[i] != 0;
    for (int k = 0; k < len; ++k) {
      const int64_t idx = base + (int64_t)k * LVPP_SLICE;
      const uint32_t c = col[idx];
      const bool cbc = (c & LVPP_COL_BC) != 0;
      const int64_t j = c & ~LVPP_COL_BC;
      double kuu = alpha * K[idx], kup = M[idx], kpu = M[idx], kpp = -D[idx];
      if (rbc) { kuu = (j == i) ? 1.0 : 0.0; kup = 0.0; }
      if (cbc) { if (!rbc) kuu = 0.0; kpu = 0.0; }
      vals[p0 + 2 * k] = kuu;
      vals[p0 + 2 * k + 1] = kup;
      vals[p1 + 2 * k] = kpu;
      vals[p1 + 2 * k + 1] = kpp;
    }
  }
}

__global__ void k_copy(int64_t n, const double* __restrict__ a, double* __restrict__ b) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    b[i] = a[i];
}
__global__ void k_flush(int64_t n, double* __restrict__ b, double v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    b[i] = v;
}

// ------------------------------------------------------------------------------------------------
// host-side dispatch
#define DISPATCH_ELEM(h, CALL)                                             \
  do {                                                                     \
    if ((h)->tdim == 2 && (h)->nld == 3) { CALL(2, 3); }                   \
    else if ((h)->tdim == 3 && (h)->nld == 4) { CALL(3, 4); }              \
    else if ((h)->tdim == 2 && (h)->nld == 6) { CALL(2, 6); }              \
    else if ((h)->tdim == 3 && (h)->nld == 10) { CALL(3, 10); }            \
    else { lvpp_set_error("unsupported element"); return LVPP_E_INVALID; } \
  } while (0)

static int gather_block(const lvpp_problem* h) {
  // per-thread accumulators live in shared memory: maxw * block * 8 bytes <= ~200 KB
  int block = 128;
  while (block > 32 && (size_t)h->maxw * block * sizeof(double) > 200 * 1024) block >>= 1;
  return block;
}

template <int NLD>
static int launch_row_gather(lvpp_problem* h, double* out) {
  const int block = gather_block(h);
  const size_t smem = (size_t)h->maxw * block * sizeof(double);
  if (smem > 220 * 1024) { lvpp_set_error("row too long for the gather kernel"); return LVPP_E_CAPACITY; }
  static bool attr_set[4] = {false, false, false, false};
  const int slot = NLD == 3 ? 0 : NLD == 4 ? 1 : NLD == 6 ? 2 : 3;
  if (!attr_set[slot] || smem > 48 * 1024) {
    CK(cudaFuncSetAttribute(k_row_gather<NLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set[slot] = true;
  }
  LAUNCH(h, k_row_gather<NLD>, lvpp_grid(h->Vown, block, 16), block, smem, h->Vown, h->inc_ptr, h->ie_ptr,
         h->ie_k, h->De, h->slice_ptr, h->rowlen, out);
  CK(cudaGetLastError());
  return 0;
}
static int row_gather(lvpp_problem* h, double* out) {
  switch (h->nld) {
    case 3: return launch_row_gather<3>(h, out);
    case 4: return launch_row_gather<4>(h, out);
    case 6: return launch_row_gather<6>(h, out);
    case 10: return launch_row_gather<10>(h, out);
  }
  return LVPP_E_INVALID;
}

int lvpp_build_constant_operators(lvpp_problem* h, const lvpp_obstacle_desc* d) {
  const int nq = h->nq, nld = h->nld, td = h->tdim;
  const size_t smem = sizeof(double) * nq * (1 + nld + nld * td + td);
  double* phi_q = nullptr;
  if (d->obstacle_kind == LVPP_OBSTACLE_ARRAY) {
    CKR(lvpp_dalloc(h, &phi_q, (size_t)h->C * nq, false));
    CK(cudaMemcpyAsync(phi_q, d->phi_obs_q, sizeof(double) * h->C * nq, cudaMemcpyHostToDevice, h->stream));
  }
  double* ve = nullptr;
  CKR(lvpp_dalloc(h, &ve, (size_t)h->C * 2 * nld, false));
  const int grid = lvpp_grid(h->C, 128, 16);
#define CALL_SETUP(TD, NL)                                                                              \
  for (int which = 0; which < 3; ++which) {                                                            \
    auto kern = k_cell_setup<TD, NL>;                                                                   \
    LAUNCH(h, kern, grid, 128, smem, which, h->C, nq, h->cells, h->coords, h->tab,                      \
           d->obstacle_kind, d->obstacle_period, d->obstacle_origin, phi_q, h->pos, h->De, h->adetJ, ve);                                               \
    CK(cudaGetLastError());                                                                             \
    if (which == 0) CKR(row_gather(h, h->K));                                                           \
    if (which == 1) CKR(row_gather(h, h->M));                                                           \
    if (which == 2) {                                                                                   \
      LAUNCH(h, k_vec_gather<NL>, lvpp_grid(h->Vown, 128, 16), 128, 0, h->Vown, h->inc_ptr, h->inc_val, \
             ve, h->bobs, h->fvec);                                                                     \
      CK(cudaGetLastError());                                                                           \
    }                                                                                                   \
  }
  DISPATCH_ELEM(h, CALL_SETUP);
#undef CALL_SETUP
  CK(cudaStreamSynchronize(h->stream));
  CKR(lvpp_dfree(h, ve));
  if (phi_q) CKR(lvpp_dfree(h, phi_q));
  return 0;
}

int lvpp_reduce_partials(lvpp_problem* h, int nvals, double* d_out) {
  LAUNCH(h, k_reduce_partials, 1, 256, 0, h->npartials, nvals, h->partials, d_out);
  CK(cudaGetLastError());
  if (h->nranks > 1) CKR(lvpp_allreduce_sum(h, d_out, nvals));
  return 0;
}

static int assemble_D(lvpp_problem* h, const double* d_x) {
  const size_t smem = sizeof(double) * h->nq * (1 + h->nld);
  const int grid = lvpp_grid(h->C, 128, 16);
  switch (h->nld) {
    case 3: LAUNCH(h, k_cell_exp<3>, grid, 128, smem, h->C, h->nq, h->cells, (const double2*)d_x, h->adetJ, h->tab, h->pos, h->De); break;
    case 4: LAUNCH(h, k_cell_exp<4>, grid, 128, smem, h->C, h->nq, h->cells, (const double2*)d_x, h->adetJ, h->tab, h->pos, h->De); break;
    case 6: LAUNCH(h, k_cell_exp<6>, grid, 128, smem, h->C, h->nq, h->cells, (const double2*)d_x, h->adetJ, h->tab, h->pos, h->De); break;
    case 10: LAUNCH(h, k_cell_exp<10>, grid, 128, smem, h->C, h->nq, h->cells, (const double2*)d_x, h->adetJ, h->tab, h->pos, h->De); break;
    default: return LVPP_E_INVALID;
  }
  CK(cudaGetLastError());
  CKR(row_gather(h, h->D));
  h->jac_valid = true;
  return 0;
}

static OpArgs op_args(lvpp_problem* h) {
  OpArgs p;
  p.Vown = h->Vown; p.slice_ptr = h->slice_ptr; p.col = h->col;
  p.K = h->K; p.M = h->M; p.D = h->D; p.bc_flag = h->bc_flag; p.bc_val = h->bc_val;
  p.alpha = h->alpha; p.v = nullptr; p.xk = (const double2*)h->xk; p.bobs = h->bobs; p.fvec = h->fvec;
  p.f = h->f; p.inv_scale = nullptr; p.skip_flag = nullptr; p.y = nullptr; p.partials = nullptr;
  p.epi = EPI_NONE; p.b = nullptr; p.binv = nullptr; p.omega = 1.0;
  return p;
}

// F(x) (owned rows) with the Jacobian at x left assembled; ||F||^2 (all ranks) lands in scal->red[0]
int lvpp_eval_residual(lvpp_problem* h, const double* d_x, double* d_F, bool want_norm, bool keep_D) {
  if (h->nranks > 1) CKR(lvpp_halo_forward_impl(h, const_cast<double*>(d_x)));
  // keep_D: d_x is the iterate D(psi) was last assembled at (lvpp_newton_begin_same_iterate) -- only alpha, f, the
  // Dirichlet values or the previous iterate changed, and D depends on none of them
  if (!(keep_D && h->jac_valid)) CKR(assemble_D(h, d_x));
  OpArgs p = op_args(h);
  p.v = (const double2*)d_x;
  p.y = (double2*)d_F;
  p.partials = want_norm ? h->partials : nullptr;
  LAUNCH(h, k_block_op<1>, h->npartials, 256, 0, p);
  CK(cudaGetLastError());
  if (want_norm) CKR(lvpp_reduce_partials(h, 1, h->scal->red));
  h->residual_evals++;
  return 0;
}

// y = J (v * *inv_scale); optional fused partial sums of (v*s).y into `partials`
int lvpp_apply_jacobian(lvpp_problem* h, const double* d_v, double* d_y, const double* inv_scale,
                        double* partials, const int* skip_flag) {
  OpArgs p = op_args(h);
  p.skip_flag = skip_flag;
  p.v = (const double2*)d_v;
  p.y = (double2*)d_y;
  p.inv_scale = inv_scale;
  p.partials = partials;
  h->fine_op_launches++;
  LAUNCH(h, k_block_op<0>, h->npartials, 256, 0, p);
  CK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// C ABI
#define CHECK_H(h)                                                       \
  do {                                                                   \
    if (!(h)) { lvpp_set_error("null handle"); return LVPP_E_INVALID; }  \
    CK(cudaSetDevice((h)->device));                                      \
  } while (0)

extern "C" int lvpp_set_alpha(lvpp_handle h, double alpha) {
  if (!h) { lvpp_set_error("null handle"); return LVPP_E_INVALID; }
  if (!(alpha > 0.0)) { lvpp_set_error("alpha must be positive"); return LVPP_E_INVALID; }
  h->alpha = alpha;
  return LVPP_OK;
}

// f.value = ... (obstacle_pg.py:74: a dolfinx Constant is read at assembly time, so a driver may change it between solves)
extern "C" int lvpp_set_forcing(lvpp_handle h, double f) {
  if (!h) { lvpp_set_error("null handle"); return LVPP_E_INVALID; }
  if (!std::isfinite(f)) { lvpp_set_error("forcing must be finite"); return LVPP_E_INVALID; }
  h->f = f;
  return LVPP_OK;
}

extern "C" int lvpp_set_previous(lvpp_handle h, const double* d_xk) {
  CHECK_H(h);
  if (!d_xk) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  CK(cudaMemcpyAsync(h->xk, d_xk, sizeof(double) * 2 * h->V, cudaMemcpyDeviceToDevice, h->stream));
  if (h->nranks > 1) CKR(lvpp_halo_forward_impl(h, h->xk));
  CKR(lvpp_sync_check_comm(h));
  return LVPP_OK;
}

extern "C" int lvpp_set_previous_host(lvpp_handle h, const double* h_xk) {
  CHECK_H(h);
  if (!h_xk) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  CK(cudaMemcpyAsync(h->xk, h_xk, sizeof(double) * 2 * h->V, cudaMemcpyHostToDevice, h->stream));
  if (h->nranks > 1) CKR(lvpp_halo_forward_impl(h, h->xk));
  CKR(lvpp_sync_check_comm(h));
  return LVPP_OK;
}

extern "C" int lvpp_assemble_residual(lvpp_handle h, const double* d_x, double* d_F, double* h_fnorm) {
  CHECK_H(h);
  if (!d_x || !d_F) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  CKR(lvpp_eval_residual(h, d_x, d_F, true));
  CK(cudaMemcpyAsync(h->red_host, h->scal->red, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CKR(lvpp_sync_check_comm(h));
  if (h_fnorm) *h_fnorm = sqrt(h->red_host[0]);
  return LVPP_OK;
}

extern "C" int lvpp_assemble_jacobian(lvpp_handle h, const double* d_x) {
  CHECK_H(h);
  if (!d_x) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  if (h->nranks > 1) CKR(lvpp_halo_forward_impl(h, const_cast<double*>(d_x)));
  CKR(assemble_D(h, d_x));
  CKR(lvpp_sync_check_comm(h));
  return LVPP_OK;
}

extern "C" int lvpp_get_jacobian_values(lvpp_handle h, double* d_values) {
  CHECK_H(h);
  if (!d_values) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  if (!h->jac_valid) { lvpp_set_error("no Jacobian assembled yet"); return LVPP_E_INVALID; }
  LAUNCH(h, k_export_values, lvpp_grid(h->Vown, 128, 16), 128, 0, h->Vown, h->rowptr, h->slice_ptr, h->rowlen,
         h->col, h->K, h->M, h->D, h->bc_flag, h->alpha, d_values);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return LVPP_OK;
}

extern "C" int lvpp_spmv(lvpp_handle h, const double* d_v, double* d_y) {
  CHECK_H(h);
  if (!d_v || !d_y) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  if (!h->jac_valid) { lvpp_set_error("no Jacobian assembled yet"); return LVPP_E_INVALID; }
  if (h->nranks > 1) CKR(lvpp_halo_forward_impl(h, const_cast<double*>(d_v)));
  CKR(lvpp_apply_jacobian(h, d_v, d_y, nullptr, nullptr, nullptr));
  CKR(lvpp_sync_check_comm(h));
  return LVPP_OK;
}

extern "C" int lvpp_observables(lvpp_handle h, const double* d_x, double* h_out6) {
  CHECK_H(h);
  if (!d_x || !h_out6) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  if (h->nranks > 1) CKR(lvpp_halo_forward_impl(h, const_cast<double*>(d_x)));
  const int nq = h->nq, nld = h->nld, td = h->tdim;
  const size_t smem = sizeof(double) * nq * (1 + nld + nld * td);
  const int grid = h->npartials;  // every block writes its partial sums
#define CALL_OBS(TD, NL)                                                                                 \
  {                                                                                                      \
    auto kern = k_observables<TD, NL>;                                                                   \
    LAUNCH(h, kern, grid, 128, smem, h->Cown, nq, h->cells, h->coords, h->tab, (const double2*)d_x,      \
           (const double2*)h->xk, h->alpha, h->f, h->npartials, h->partials);                            \
  }
  DISPATCH_ELEM(h, CALL_OBS);
#undef CALL_OBS
  CK(cudaGetLastError());
  CKR(lvpp_reduce_partials(h, 6, h->scal->red));
  CK(cudaMemcpyAsync(h->red_host, h->scal->red, 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CKR(lvpp_sync_check_comm(h));
  for (int i = 0; i < 6; ++i) h_out6[i] = h->red_host[i];
  return LVPP_OK;
}

static int ensure_flush(lvpp_problem* h) {
  if (!h->flush) {
    h->flush_bytes = (size_t)256 << 20;  // 256 MiB > 126 MB L2
    CK(cudaMalloc((void**)&h->flush, h->flush_bytes));
  }
  return 0;
}

extern "C" int lvpp_time_spmv(lvpp_handle h, const double* d_v, double* d_y, int32_t reps, int32_t flush_l2,
                              double* h_ms) {
  CHECK_H(h);
  if (!d_v || !d_y || !h_ms || reps < 1) { lvpp_set_error("bad argument"); return LVPP_E_INVALID; }
  if (!h->jac_valid) { lvpp_set_error("no Jacobian assembled yet"); return LVPP_E_INVALID; }
  if (flush_l2) CKR(ensure_flush(h));
  double total = 0.0;
  for (int r = 0; r < reps; ++r) {
    if (flush_l2) {
      LAUNCH(h, k_flush, LVPP_NUM_SMS * 8, 256, 0, (int64_t)(h->flush_bytes / 8), h->flush, (double)r);
      CK(cudaGetLastError());
    }
    CK(cudaEventRecord(h->ev0, h->stream));
    CKR(lvpp_apply_jacobian(h, d_v, d_y, nullptr, nullptr, nullptr));
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    total += ms;
  }
  *h_ms = total / reps;
  h->last_spmv_ms = *h_ms;
  return LVPP_OK;
}

extern "C" int lvpp_time_assembly(lvpp_handle h, const double* d_x, double* d_F, int32_t reps, double* h_ms_cells,
                                  double* h_ms_gather, double* h_ms_residual) {
  CHECK_H(h);
  if (!d_x || !d_F || reps < 1) { lvpp_set_error("bad argument"); return LVPP_E_INVALID; }
  double t[3] = {0, 0, 0};
  const size_t smem = sizeof(double) * h->nq * (1 + h->nld);
  const int gridc = lvpp_grid(h->C, 128, 16);
  for (int r = 0; r < reps; ++r) {
    float ms = 0.f;
    CK(cudaEventRecord(h->ev0, h->stream));
    switch (h->nld) {
      case 3: LAUNCH(h, k_cell_exp<3>, gridc, 128, smem, h->C, h->nq, h->cells, (const double2*)d_x, h->adetJ, h->tab, h->pos, h->De); break;
      case 4: LAUNCH(h, k_cell_exp<4>, gridc, 128, smem, h->C, h->nq, h->cells, (const double2*)d_x, h->adetJ, h->tab, h->pos, h->De); break;
      case 6: LAUNCH(h, k_cell_exp<6>, gridc, 128, smem, h->C, h->nq, h->cells, (const double2*)d_x, h->adetJ, h->tab, h->pos, h->De); break;
      case 10: LAUNCH(h, k_cell_exp<10>, gridc, 128, smem, h->C, h->nq, h->cells, (const double2*)d_x, h->adetJ, h->tab, h->pos, h->De); break;
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaEventSynchronize(h->ev1));
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    t[0] += ms;
    CK(cudaEventRecord(h->ev0, h->stream));
    CKR(row_gather(h, h->D));
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaEventSynchronize(h->ev1));
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    t[1] += ms;
    h->jac_valid = true;
    OpArgs p = op_args(h);
    p.v = (const double2*)d_x;
    p.y = (double2*)d_F;
    p.partials = h->partials;
    CK(cudaEventRecord(h->ev0, h->stream));
    LAUNCH(h, k_block_op<1>, h->npartials, 256, 0, p);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaEventSynchronize(h->ev1));
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    t[2] += ms;
  }
  if (h_ms_cells) *h_ms_cells = t[0] / reps;
  if (h_ms_gather) *h_ms_gather = t[1] / reps;
  if (h_ms_residual) *h_ms_residual = t[2] / reps;
  return LVPP_OK;
}
