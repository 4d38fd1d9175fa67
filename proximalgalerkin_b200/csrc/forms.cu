// Generic mixed-form engine for the LVPP formulations of SURVEY.md section 8a rows a13-a18 (see
// include/lvpp_b200.h, "lvpp_form_*"): what dolfinx + FFCx + PETSc do for
//   examples/06_gradient_constraints/gradient_constraint_dolfinx.py:100-132   (LVPP_FORM_GRADIENT)
//   examples/04_multiphase/multiphase_dolfinx.py:64-147                       (LVPP_FORM_MULTIPHASE)
//   examples/02_signorini/signorini_dolfinx.py:244-291                        (LVPP_FORM_SIGNORINI)
// restructured for the GPU:
//   * one thread per integration entity evaluates the element residual vector and a compact form of
//     the element Jacobian (quadrature tables staged in shared memory, pointwise latent maps
//     psi/sqrt(1+|psi|^2), softmax, exp evaluated once per quadrature point);
//   * the scatter-add of MatSetValuesLocal / VecSetValuesLocal is inverted into gathers through
//     precomputed sorted contribution lists (one thread per CSR entry / per row; fixed summation order,
//     no atomics), with the Dirichlet rows/columns of assemble_matrix(bcs) applied on the fly;
//   * lifting: all three forms are linear in the constrained field, so apply_lifting(x0 = x, scale -1)
//     (src/lvpp/problem.py:59-67) equals evaluating the residual at x with its Dirichlet entries
//     replaced by g; Dirichlet rows are x - g;
//   * Krylov: right-preconditioned restarted GMRES, dof-block Jacobi (dense inverse of the diagonal
//     block of every mesh node's dofs); Newton: SNES newtonls, line search none or bt.
#include <cub/cub.cuh>

#include <cmath>
#include <cstring>

#include "gmres_kernels.cuh"
#include "lvpp_internal.cuh"

#define FORM_MAX_BLOCK 16
#define FORM_MAX_INTEGRALS 2
#define FORM_MAX_PARAMS 16

struct IntegralDev {
  int64_t E = 0;
  int nld = 0, nv = 0, nq = 0;
  int32_t* dofs = nullptr;
  int32_t* verts = nullptr;
  double* tab = nullptr;      // w[nq] | tab_a | dtab_a | tab_b
  int tab_len = 0, off_a = 0, off_da = 0, off_b = 0;
  int cd_stride = 0;          // doubles of compact element-Jacobian data per entity
  double* cd = nullptr;       // [E * cd_stride]
  double* be = nullptr;       // [E * nld] element residual vectors
  int64_t mat_off = 0;        // first flat (entity, i, j) index of this integral
  int64_t vec_off = 0;        // first flat (entity, i) index
};

struct lvpp_form_problem {
  int device = 0;
  cudaStream_t stream = nullptr;
  int form = 0, gdim = 0;
  int64_t n = 0, npad = 0, nnz = 0, nverts = 0;
  int64_t launches = 0, device_bytes = 0;
  std::vector<std::pair<void*, size_t>> allocs;
  int64_t* indptr = nullptr;
  int32_t* indices = nullptr;
  int32_t* nnz_row = nullptr;
  double* vals = nullptr;
  uint8_t* bc_flag = nullptr;
  double* bc_val = nullptr;
  int64_t num_bc = 0;
  int64_t* bc_dofs = nullptr;
  double* coords = nullptr;
  int nint = 0;
  IntegralDev itg[FORM_MAX_INTEGRALS];
  int64_t* mptr = nullptr;   // [nnz + 1] contribution segments per CSR entry
  uint32_t* msrc = nullptr;  // flat (entity, i, j) indices, ascending inside a segment
  int64_t* vptr = nullptr;   // [n + 1]
  uint32_t* vsrc = nullptr;
  double params[FORM_MAX_PARAMS] = {0};
  double *aux0 = nullptr, *aux1 = nullptr, *coef0 = nullptr, *coef1 = nullptr;
  double *xt = nullptr, *F = nullptr, *y = nullptr, *w = nullptr, *G = nullptr, *Jy = nullptr, *z = nullptr;
  bool jac_valid = false;
  // preconditioner
  int64_t nblk = 0;
  int64_t* blk_ptr = nullptr;
  int32_t* blk_dofs = nullptr;
  int64_t* blk_inv_ptr = nullptr;
  double* blk_inv = nullptr;
  // GMRES
  int gm_restart = 0;
  double* gm_V = nullptr;
  double *gm_h = nullptr, *gm_h_host = nullptr, *gm_part = nullptr;
  // device-resident GMRES recurrence (gmres_kernels.cuh: GmState), as in multigrid.cu
  GmState* gm_state = nullptr;       // device
  GmState* gm_state_host = nullptr;  // pinned [GM_RING + 2]
  double *gm_H = nullptr, *gm_cs = nullptr, *gm_sn = nullptr, *gm_g = nullptr, *gm_yv = nullptr, *gm_red = nullptr;
  cudaEvent_t gm_ev[GM_RING] = {nullptr};
  int npartials = 0;
  double* red_host = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int64_t krylov_its = 0, newton_steps = 0, residual_evals = 0, spmv_launches = 0;
  double t_assembly_ms = 0.0, t_krylov_ms = 0.0;
  int spmv_tpr = 8;
};

template <class T>
static int fdalloc(lvpp_form_problem* h, T** p, size_t n, bool zero = true) {
  size_t bytes = (n ? n : 1) * sizeof(T);
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, bytes);
  if (e != cudaSuccess) { lvpp_set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e)); return LVPP_E_CUDA; }
  if (zero) {
    e = cudaMemsetAsync(q, 0, bytes, h->stream);
    if (e != cudaSuccess) { lvpp_set_error("cudaMemset failed: %s", cudaGetErrorString(e)); return LVPP_E_CUDA; }
  }
  h->allocs.push_back(std::make_pair(q, bytes));
  h->device_bytes += (int64_t)bytes;
  *p = (T*)q;
  return 0;
}
static int fdfree(lvpp_form_problem* h, void* p) {
  for (size_t i = 0; i < h->allocs.size(); ++i)
    if (h->allocs[i].first == p) {
      CK(cudaStreamSynchronize(h->stream));
      CK(cudaFree(p));
      h->device_bytes -= (int64_t)h->allocs[i].second;
      h->allocs.erase(h->allocs.begin() + i);
      return 0;
    }
  return 0;
}
template <class T>
static int fupload(lvpp_form_problem* h, T** p, const T* src, size_t n) {
  CKR(fdalloc(h, p, n, false));
  if (n) CK(cudaMemcpyAsync(*p, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  return 0;
}

// ================================================================================================
// element kernels
struct ElemArgs {
  int64_t E;
  int nq;
  const int32_t* dofs;
  const int32_t* verts;
  const double* coords;
  const double* tab;
  int off_a, off_da, off_b, tab_len;
  const double *xt, *aux0, *aux1, *coef0, *coef1;
  double p[FORM_MAX_PARAMS];
  double* cd;
  double* be;
};

__device__ __forceinline__ int sym3(int a, int b) {  // packed upper triangle of a symmetric 3x3
  return a <= b ? a * 3 - (a * (a - 1)) / 2 + (b - a) : b * 3 - (b * (b - 1)) / 2 + (a - b);
}
__device__ __forceinline__ int sym4(int a, int b) {
  return a <= b ? a * 4 - (a * (a - 1)) / 2 + (b - a) : b * 4 - (b * (b - 1)) / 2 + (a - b);
}
__device__ __forceinline__ int sym6(int a, int b) {
  return a <= b ? a * 6 - (a * (a - 1)) / 2 + (b - a) : b * 6 - (b * (b - 1)) / 2 + (a - b);
}

// ---- gradient constraint: compact data = K[21] (P2 stiffness, packed), B[6][3][2], D[6 vertex pairs][3 (xx,xy,yy)]
#define GRAD_CD (21 + 36 + 18)
__global__ void __launch_bounds__(128) k_elem_gradient(ElemArgs a) {
  extern __shared__ double s_tab[];
  stage_tables(s_tab, a.tab, a.tab_len);
  const double* s_w = s_tab;
  const double* s_p2 = s_tab + a.off_a;
  const double* s_d2 = s_tab + a.off_da;
  const double* s_p1 = s_tab + a.off_b;
  const double alpha = a.p[0];
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < a.E; c += (int64_t)gridDim.x * blockDim.x) {
    int32_t d[12], vt[3];
#pragma unroll
    for (int i = 0; i < 12; ++i) d[i] = a.dofs[c * 12 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) vt[i] = a.verts[c * 3 + i];
    double Jinv[2][2];
    const double adet = cell_geometry<2>(a.coords, vt, Jinv);
    double u[6], ph[6], fh[6], ps[3][2], p0[3][2];
#pragma unroll
    for (int i = 0; i < 6; ++i) { u[i] = a.xt[d[i]]; ph[i] = a.coef0[d[i]]; fh[i] = a.coef1[d[i]]; }
#pragma unroll
    for (int v = 0; v < 3; ++v)
#pragma unroll
      for (int g = 0; g < 2; ++g) { ps[v][g] = a.xt[d[6 + 2 * v + g]]; p0[v][g] = a.aux0[d[6 + 2 * v + g]]; }
    double Fu[6], Fp[3][2], K[21], B[6][3][2], D[6][3];
#pragma unroll
    for (int i = 0; i < 6; ++i) Fu[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 21; ++i) K[i] = 0.0;
#pragma unroll
    for (int v = 0; v < 3; ++v) { Fp[v][0] = Fp[v][1] = 0.0; }
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int v = 0; v < 3; ++v) B[i][v][0] = B[i][v][1] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) D[i][0] = D[i][1] = D[i][2] = 0.0;
    for (int q = 0; q < a.nq; ++q) {
      const double w = s_w[q] * adet;
      double g2[6][2], gu[2] = {0.0, 0.0}, fq = 0.0, pq = 0.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const double dx = s_d2[(q * 6 + i) * 2], dy = s_d2[(q * 6 + i) * 2 + 1];
        g2[i][0] = dx * Jinv[0][0] + dy * Jinv[1][0];
        g2[i][1] = dx * Jinv[0][1] + dy * Jinv[1][1];
        gu[0] += u[i] * g2[i][0];
        gu[1] += u[i] * g2[i][1];
        fq += fh[i] * s_p2[q * 6 + i];
        pq += ph[i] * s_p2[q * 6 + i];
      }
      double psq[2] = {0.0, 0.0}, p0q[2] = {0.0, 0.0};
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        const double l = s_p1[q * 3 + v];
        psq[0] += l * ps[v][0]; psq[1] += l * ps[v][1];
        p0q[0] += l * p0[v][0]; p0q[1] += l * p0[v][1];
      }
      const double s = sqrt(1.0 + psq[0] * psq[0] + psq[1] * psq[1]);
      const double fl0 = alpha * gu[0] + psq[0] - p0q[0], fl1 = alpha * gu[1] + psq[1] - p0q[1];
#pragma unroll
      for (int i = 0; i < 6; ++i)
        Fu[i] += w * (fl0 * g2[i][0] + fl1 * g2[i][1] - alpha * fq * s_p2[q * 6 + i]);
      const double r0 = gu[0] - pq / s * psq[0], r1 = gu[1] - pq / s * psq[1];
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        Fp[v][0] += w * r0 * s_p1[q * 3 + v];
        Fp[v][1] += w * r1 * s_p1[q * 3 + v];
      }
      int k = 0;
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = i; j < 6; ++j, ++k) K[k] += w * (g2[i][0] * g2[j][0] + g2[i][1] * g2[j][1]);
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int v = 0; v < 3; ++v) {
          B[i][v][0] += w * g2[i][0] * s_p1[q * 3 + v];
          B[i][v][1] += w * g2[i][1] * s_p1[q * 3 + v];
        }
      // d/dpsi [phi psi / s] = phi (I / s - psi psi^T / s^3)
      const double is = pq / s, is3 = pq / (s * s * s);
      const double t00 = is - is3 * psq[0] * psq[0], t01 = -is3 * psq[0] * psq[1], t11 = is - is3 * psq[1] * psq[1];
      k = 0;
#pragma unroll
      for (int v = 0; v < 3; ++v)
#pragma unroll
        for (int x = v; x < 3; ++x, ++k) {
          const double m = w * s_p1[q * 3 + v] * s_p1[q * 3 + x];
          D[k][0] += m * t00; D[k][1] += m * t01; D[k][2] += m * t11;
        }
    }
    double* be = a.be + c * 12;
#pragma unroll
    for (int i = 0; i < 6; ++i) be[i] = Fu[i];
#pragma unroll
    for (int v = 0; v < 3; ++v) { be[6 + 2 * v] = Fp[v][0]; be[6 + 2 * v + 1] = Fp[v][1]; }
    double* cd = a.cd + c * GRAD_CD;
#pragma unroll
    for (int i = 0; i < 21; ++i) cd[i] = K[i];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int v = 0; v < 3; ++v) { cd[21 + (i * 3 + v) * 2] = B[i][v][0]; cd[21 + (i * 3 + v) * 2 + 1] = B[i][v][1]; }
#pragma unroll
    for (int i = 0; i < 6; ++i) { cd[57 + i * 3] = D[i][0]; cd[57 + i * 3 + 1] = D[i][1]; cd[57 + i * 3 + 2] = D[i][2]; }
  }
}
__device__ __forceinline__ double entry_gradient(const double* cd, const double* p, int i, int j) {
  if (i < 6 && j < 6) return p[0] * cd[sym6(i, j)];
  if (i < 6) return cd[21 + (i * 3 + (j - 6) / 2) * 2 + ((j - 6) & 1)];
  if (j < 6) return cd[21 + (j * 3 + (i - 6) / 2) * 2 + ((i - 6) & 1)];
  const int v = (i - 6) >> 1, g = (i - 6) & 1, x = (j - 6) >> 1, hh = (j - 6) & 1;
  return -cd[57 + sym3(v, x) * 3 + g + hh];
}

// ---- multiphase: compact data = M[6], K[6] (packed 3x3), N[6 vertex pairs][10 species pairs], eps^2
#define MP_CD (6 + 6 + 60 + 1)
__global__ void __launch_bounds__(128) k_elem_multiphase(ElemArgs a) {
  extern __shared__ double s_tab[];
  stage_tables(s_tab, a.tab, a.tab_len);
  const double* s_w = s_tab;
  const double* s_p1 = s_tab + a.off_a;
  const double alpha = a.p[0], tau = a.p[1], eps0 = a.p[2], hscale = a.p[3];
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < a.E; c += (int64_t)gridDim.x * blockDim.x) {
    int32_t vt[3];
#pragma unroll
    for (int v = 0; v < 3; ++v) vt[v] = a.verts[c * 3 + v];
    const int32_t* dl = a.dofs + c * 36;
    double Jinv[2][2];
    const double adet = cell_geometry<2>(a.coords, vt, Jinv);
    // physical P1 gradients, circumradius
    const double gr[3][2] = {{-Jinv[0][0] - Jinv[1][0], -Jinv[0][1] - Jinv[1][1]}, {Jinv[0][0], Jinv[0][1]}, {Jinv[1][0], Jinv[1][1]}};
    double xy[3][2];
#pragma unroll
    for (int v = 0; v < 3; ++v) { xy[v][0] = a.coords[(int64_t)vt[v] * 2]; xy[v][1] = a.coords[(int64_t)vt[v] * 2 + 1]; }
    const double la = hypot(xy[1][0] - xy[2][0], xy[1][1] - xy[2][1]), lb = hypot(xy[0][0] - xy[2][0], xy[0][1] - xy[2][1]),
                 lc = hypot(xy[0][0] - xy[1][0], xy[0][1] - xy[1][1]);
    const double eps = hscale * (la * lb * lc / (4.0 * (0.5 * adet)));
    const double e2 = eps * eps;
    // pass 1 over the quadrature points: everything that depends on psi (softmax) plus the mass matrix.
    // Only psi is live here; the terms that are linear in (u, z, psi) are applied through M and K afterwards
    // (same quadrature sums, reassociated) so that the accumulators fit the register file.
    double ps[3][4];
#pragma unroll
    for (int v = 0; v < 3; ++v)
#pragma unroll
      for (int m = 0; m < 4; ++m) ps[v][m] = a.xt[dl[v * 12 + 8 + m]];
    double M[6], N[6][10], S[3][4];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      M[k] = 0.0;
#pragma unroll
      for (int t = 0; t < 10; ++t) N[k][t] = 0.0;
    }
#pragma unroll
    for (int v = 0; v < 3; ++v)
#pragma unroll
      for (int m = 0; m < 4; ++m) S[v][m] = 0.0;
    for (int q = 0; q < a.nq; ++q) {
      const double w = s_w[q] * adet;
      const double l[3] = {s_p1[q * 3], s_p1[q * 3 + 1], s_p1[q * 3 + 2]};
      double sm[4], esum = 0.0;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        sm[m] = exp(l[0] * ps[0][m] + l[1] * ps[1][m] + l[2] * ps[2][m]);
        esum += sm[m];
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) sm[m] = sm[m] / esum;
#pragma unroll
      for (int v = 0; v < 3; ++v)
#pragma unroll
        for (int m = 0; m < 4; ++m) S[v][m] += w * sm[m] * l[v];
      int k = 0;
#pragma unroll
      for (int v = 0; v < 3; ++v)
#pragma unroll
        for (int x = v; x < 3; ++x, ++k) {
          const double mm = w * l[v] * l[x];
          M[k] += mm;
          int t = 0;
#pragma unroll
          for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int nn = m; nn < 4; ++nn, ++t) N[k][t] += mm * ((m == nn ? sm[m] : 0.0) - sm[m] * sm[nn]);
        }
    }
    double K[6];
    {
      const double area = 0.5 * adet;  // constant gradients: the rule integrates 1 exactly
      int k = 0;
#pragma unroll
      for (int v = 0; v < 3; ++v)
#pragma unroll
        for (int x = v; x < 3; ++x, ++k) K[k] = area * (gr[v][0] * gr[x][0] + gr[v][1] * gr[x][1]);
    }
    double* cd = a.cd + c * MP_CD;
#pragma unroll
    for (int k = 0; k < 6; ++k) { cd[k] = M[k]; cd[6 + k] = K[k]; }
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
      for (int t = 0; t < 10; ++t) cd[12 + k * 10 + t] = N[k][t];
    cd[72] = e2;
    // pass 2: residual rows through M and K
    double* be = a.be + c * 36;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      double u[3], z[3], dp[3], du[3], pm[3];
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        u[x] = a.xt[dl[x * 12 + m]];
        z[x] = a.xt[dl[x * 12 + 4 + m]];
        pm[x] = ps[x][m];
        dp[x] = pm[x] - a.aux0[dl[x * 12 + 8 + m]];
        du[x] = u[x] - a.aux1[dl[x * 12 + m]];
      }
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        double rz = 0.0, ru = 0.0, rp = 0.0, msum = 0.0;
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          const double mvx = M[sym3(v, x)], kvx = K[sym3(v, x)];
          msum += mvx;
          rz += mvx * (alpha * z[x] - 2.0 * alpha * u[x] + dp[x]) + alpha * e2 * kvx * u[x];
          ru += mvx * du[x] - tau * kvx * z[x];
          rp += mvx * (u[x] - eps0 * pm[x]);
        }
        be[v * 12 + m] = ru;                       // EQ2, test v
        be[v * 12 + 4 + m] = rz - alpha * msum;    // EQ1, test y
        be[v * 12 + 8 + m] = rp - S[v][m];         // EQ3, test w
      }
    }
  }
}
__device__ __forceinline__ double entry_multiphase(const double* cd, const double* p, int i, int j) {
  const int v = i / 12, si = (i % 12) >> 2, m = i & 3, x = j / 12, sj = (j % 12) >> 2, nn = j & 3;
  const int k = sym3(v, x);
  const double alpha = p[0], tau = p[1], eps0 = p[2];
  const double Mk = cd[k], Kk = cd[6 + k];
  if (si == 2 && sj == 2) return -cd[12 + k * 10 + sym4(m, nn)] - (m == nn ? eps0 * Mk : 0.0);
  if (m != nn) return 0.0;
  if (si == 0) return sj == 0 ? Mk : (sj == 1 ? -tau * Kk : 0.0);                     // EQ2 (test v, u rows)
  if (si == 1) return sj == 0 ? alpha * cd[72] * Kk - 2.0 * alpha * Mk : (sj == 1 ? alpha * Mk : Mk);  // EQ1 (test y)
  return sj == 0 ? Mk : 0.0;                                                           // EQ3 (test w), u column
}

// ---- Signorini, integral 0 (tetrahedra): compact data = elasticity stiffness, full 12 x 12 (times alpha at use)
#define SIG_CD0 144
__global__ void __launch_bounds__(128) k_elem_elasticity(ElemArgs a) {
  const double alpha = a.p[0], mu = a.p[1], lm = a.p[2];
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < a.E; c += (int64_t)gridDim.x * blockDim.x) {
    int32_t vt[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) vt[v] = a.verts[c * 4 + v];
    double Jinv[3][3];
    const double vol = cell_geometry<3>(a.coords, vt, Jinv) / 6.0;
    double g[4][3];
#pragma unroll
    for (int dd = 0; dd < 3; ++dd) {
      g[0][dd] = -Jinv[0][dd] - Jinv[1][dd] - Jinv[2][dd];
      g[1][dd] = Jinv[0][dd]; g[2][dd] = Jinv[1][dd]; g[3][dd] = Jinv[2][dd];
    }
    double ul[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) ul[i] = a.xt[a.dofs[c * 12 + i]];
    double* cd = a.cd + c * SIG_CD0;
    double r[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) r[i] = 0.0;
#pragma unroll
    for (int va = 0; va < 4; ++va)
#pragma unroll
      for (int vb = 0; vb < 4; ++vb) {
        const double gg = g[va][0] * g[vb][0] + g[va][1] * g[vb][1] + g[va][2] * g[vb][2];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const double e = vol * (mu * ((i == j ? gg : 0.0) + g[va][j] * g[vb][i]) + lm * g[va][i] * g[vb][j]);
            cd[(va * 3 + i) * 12 + vb * 3 + j] = e;
            r[va * 3 + i] += alpha * e * ul[vb * 3 + j];
          }
      }
#pragma unroll
    for (int i = 0; i < 12; ++i) a.be[c * 12 + i] = r[i];
  }
}
// ---- Signorini, integral 1 (contact facets): compact data = M[6], D[6] (packed 3x3)
#define SIG_CD1 12
__global__ void __launch_bounds__(128) k_elem_contact(ElemArgs a) {
  extern __shared__ double s_tab[];
  stage_tables(s_tab, a.tab, a.tab_len);
  const double* s_w = s_tab;
  const double* s_p1 = s_tab + a.off_a;
  const double gap = a.p[3], n0 = a.p[4], n1 = a.p[5], n2 = a.p[6];
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < a.E; c += (int64_t)gridDim.x * blockDim.x) {
    double X[3][3];
#pragma unroll
    for (int v = 0; v < 3; ++v)
#pragma unroll
      for (int dd = 0; dd < 3; ++dd) X[v][dd] = a.coords[(int64_t)a.verts[c * 3 + v] * 3 + dd];
    const double e1[3] = {X[1][0] - X[0][0], X[1][1] - X[0][1], X[1][2] - X[0][2]};
    const double e2[3] = {X[2][0] - X[0][0], X[2][1] - X[0][1], X[2][2] - X[0][2]};
    const double cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
    const double area2 = sqrt(cx * cx + cy * cy + cz * cz);
    double un_v[3], p[3], pk[3];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      un_v[v] = a.xt[a.dofs[c * 12 + v * 3]] * n0 + a.xt[a.dofs[c * 12 + v * 3 + 1]] * n1 + a.xt[a.dofs[c * 12 + v * 3 + 2]] * n2;
      p[v] = a.xt[a.dofs[c * 12 + 9 + v]];
      pk[v] = a.aux0[a.dofs[c * 12 + 9 + v]];
    }
    double M[6] = {0, 0, 0, 0, 0, 0}, D[6] = {0, 0, 0, 0, 0, 0}, Rn[3] = {0, 0, 0}, Rp[3] = {0, 0, 0};
    for (int q = 0; q < a.nq; ++q) {
      const double w = s_w[q] * area2;
      const double l[3] = {s_p1[q * 3], s_p1[q * 3 + 1], s_p1[q * 3 + 2]};
      const double un = l[0] * un_v[0] + l[1] * un_v[1] + l[2] * un_v[2];
      const double pq = l[0] * p[0] + l[1] * p[1] + l[2] * p[2];
      const double pkq = l[0] * pk[0] + l[1] * pk[1] + l[2] * pk[2];
      const double gq = l[0] * X[0][2] + l[1] * X[1][2] + l[2] * X[2][2] - gap;
      const double ex = exp(pq);
      int k = 0;
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        Rn[v] += w * (pq - pkq) * l[v];
        Rp[v] += w * (un + ex - gq) * l[v];
#pragma unroll
        for (int x = v; x < 3; ++x, ++k) { M[k] += w * l[v] * l[x]; D[k] += w * ex * l[v] * l[x]; }
      }
    }
    double* be = a.be + c * 12;
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      be[v * 3] = -Rn[v] * n0; be[v * 3 + 1] = -Rn[v] * n1; be[v * 3 + 2] = -Rn[v] * n2;
      be[9 + v] = Rp[v];
    }
    double* cd = a.cd + c * SIG_CD1;
#pragma unroll
    for (int k = 0; k < 6; ++k) { cd[k] = M[k]; cd[6 + k] = D[k]; }
  }
}
__device__ __forceinline__ double entry_signorini(int blk, const double* cd, const double* p, int i, int j) {
  if (blk == 0) return p[0] * cd[i * 12 + j];
  const double ng[3] = {p[4], p[5], p[6]};
  if (i < 9 && j < 9) return 0.0;
  if (i < 9) return -cd[sym3(i / 3, j - 9)] * ng[i % 3];
  if (j < 9) return cd[sym3(i - 9, j / 3)] * ng[j % 3];
  return cd[6 + sym3(i - 9, j - 9)];
}

// ================================================================================================
// gathers
struct GatherArgs {
  int form, nint;
  const double* cd[FORM_MAX_INTEGRALS];
  const double* be[FORM_MAX_INTEGRALS];
  int cd_stride[FORM_MAX_INTEGRALS], nld[FORM_MAX_INTEGRALS];
  int64_t mat_off[FORM_MAX_INTEGRALS + 1], vec_off[FORM_MAX_INTEGRALS + 1];
  double p[FORM_MAX_PARAMS];
};

// one thread per CSR entry: sums its element contributions in ascending (entity, i, j) order and applies
// the Dirichlet rows / columns of assemble_matrix(bcs): zero with unit diagonal
__global__ void __launch_bounds__(256) k_gather_matrix(int64_t nnz, GatherArgs g, const int64_t* __restrict__ mptr,
                                                        const uint32_t* __restrict__ msrc, const int32_t* __restrict__ nnz_row,
                                                        const int32_t* __restrict__ indices, const uint8_t* __restrict__ bc,
                                                        double* __restrict__ vals) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) {
    const int32_t r = nnz_row[k], cidx = indices[k];
    if (bc[r] || bc[cidx]) { vals[k] = (r == cidx && bc[r]) ? 1.0 : 0.0; continue; }
    double s = 0.0;
    const int64_t p1 = mptr[k + 1];
    for (int64_t p = mptr[k]; p < p1; ++p) {
      int64_t f = msrc[p];
      const int blk = (g.nint > 1 && f >= g.mat_off[1]) ? 1 : 0;
      f -= g.mat_off[blk];
      const int n = g.nld[blk], n2 = n * n;
      const int64_t e = f / n2;
      const int ij = (int)(f - e * n2);
      const int i = ij / n, j = ij - i * n;
      const double* cd = g.cd[blk] + e * g.cd_stride[blk];
      double v;
      if (g.form == LVPP_FORM_GRADIENT) v = entry_gradient(cd, g.p, i, j);
      else if (g.form == LVPP_FORM_MULTIPHASE) v = entry_multiphase(cd, g.p, i, j);
      else v = entry_signorini(blk, cd, g.p, i, j);
      s += v;
    }
    vals[k] = s;
  }
}
// one thread per row: residual = sum of element vector entries; Dirichlet rows x - g (set_bc with scale -1)
__global__ void __launch_bounds__(256) k_gather_vector(int64_t n, GatherArgs g, const int64_t* __restrict__ vptr,
                                                        const uint32_t* __restrict__ vsrc, const uint8_t* __restrict__ bc,
                                                        const double* __restrict__ bcv, const double* __restrict__ x,
                                                        double* __restrict__ F, int nparts, double* __restrict__ partials) {
  __shared__ double s_red[32];
  double part = 0.0;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    if (bc[r]) {
      s = x[r] - bcv[r];
    } else {
      const int64_t p1 = vptr[r + 1];
      for (int64_t p = vptr[r]; p < p1; ++p) {
        int64_t f = vsrc[p];
        const int blk = (g.nint > 1 && f >= g.vec_off[1]) ? 1 : 0;
        s += g.be[blk][f - g.vec_off[blk]];
      }
    }
    F[r] = s;
    part += s * s;
  }
  const double rr = lvpp_block_sum<256>(part, s_red);
  if (threadIdx.x == 0) partials[blockIdx.x] = rr;
}
__global__ void k_apply_bc_values(int64_t n, const double* __restrict__ x, const uint8_t* __restrict__ bc,
                                  const double* __restrict__ bcv, double* __restrict__ xt) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    xt[i] = bc[i] ? bcv[i] : x[i];
}
__global__ void k_nnz_rows(int64_t n, const int64_t* __restrict__ indptr, int32_t* __restrict__ nnz_row) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
    for (int64_t k = indptr[r]; k < indptr[r + 1]; ++k) nnz_row[k] = (int32_t)r;
}
__global__ void k_iota_u32(int64_t n, uint32_t base, uint32_t* __restrict__ v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    v[i] = base + (uint32_t)i;
}
__global__ void k_i32_to_i64(int64_t n, const int32_t* __restrict__ a, int64_t* __restrict__ b) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) b[i] = a[i];
}
// segment pointers of a sorted key array: ptr[k] = first position with key >= k (k = 0..nkeys)
__global__ void k_segment_ptr(int64_t nkeys, int64_t m, const int64_t* __restrict__ skeys, int64_t* __restrict__ ptr) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k <= nkeys; k += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = 0, hi = m;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (skeys[mid] < k) lo = mid + 1; else hi = mid;
    }
    ptr[k] = lo;
  }
}

// ================================================================================================
// CSR SpMV: TPR threads per row (power of two <= 32), fp64; fused partial sum of v . y when requested
template <int TPR>
__global__ void __launch_bounds__(256) k_csr_spmv(int64_t n, const int64_t* __restrict__ indptr,
                                                   const int32_t* __restrict__ indices, const double* __restrict__ vals,
                                                   const double* __restrict__ v, double* __restrict__ y) {
  const int lane = threadIdx.x & (TPR - 1);
  for (int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / TPR; r < n;
       r += ((int64_t)gridDim.x * blockDim.x) / TPR) {
    double s = 0.0;
    const int64_t k1 = indptr[r + 1];
    for (int64_t k = indptr[r] + lane; k < k1; k += TPR) s += vals[k] * __ldg(&v[indices[k]]);
#pragma unroll
    for (int o = TPR >> 1; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, TPR);
    if (lane == 0) y[r] = s;
  }
}

// ================================================================================================
// dof-block Jacobi: dense inverse of every diagonal block (Gauss-Jordan with partial pivoting, one thread per block)
__global__ void __launch_bounds__(64) k_block_inverse(int64_t nblk, const int64_t* __restrict__ bptr,
                                                       const int32_t* __restrict__ bdofs, const int64_t* __restrict__ iptr,
                                                       const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                       const double* __restrict__ vals, double* __restrict__ inv,
                                                       int* __restrict__ err) {
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < nblk; b += (int64_t)gridDim.x * blockDim.x) {
    const int64_t d0 = bptr[b];
    const int m = (int)(bptr[b + 1] - d0);
    double A[FORM_MAX_BLOCK * FORM_MAX_BLOCK], I[FORM_MAX_BLOCK * FORM_MAX_BLOCK];
    for (int i = 0; i < m; ++i) {
      const int32_t r = bdofs[d0 + i];
      for (int j = 0; j < m; ++j) {
        const int32_t cc = bdofs[d0 + j];
        int64_t lo = indptr[r], hi = indptr[r + 1];
        while (lo < hi) {
          const int64_t mid = (lo + hi) >> 1;
          if (indices[mid] < cc) lo = mid + 1; else hi = mid;
        }
        A[i * m + j] = (lo < indptr[r + 1] && indices[lo] == cc) ? vals[lo] : 0.0;
        I[i * m + j] = i == j ? 1.0 : 0.0;
      }
    }
    for (int k = 0; k < m; ++k) {
      int piv = k;
      double best = fabs(A[k * m + k]);
      for (int r = k + 1; r < m; ++r)
        if (fabs(A[r * m + k]) > best) { best = fabs(A[r * m + k]); piv = r; }
      if (!(best > 0.0)) { *err = 1; best = 1.0; A[k * m + k] = 1.0; piv = k; }
      if (piv != k)
        for (int cc = 0; cc < m; ++cc) {
          double t = A[k * m + cc]; A[k * m + cc] = A[piv * m + cc]; A[piv * m + cc] = t;
          t = I[k * m + cc]; I[k * m + cc] = I[piv * m + cc]; I[piv * m + cc] = t;
        }
      const double ip = 1.0 / A[k * m + k];
      for (int cc = 0; cc < m; ++cc) { A[k * m + cc] *= ip; I[k * m + cc] *= ip; }
      for (int r = 0; r < m; ++r) {
        if (r == k) continue;
        const double f = A[r * m + k];
        if (f != 0.0)
          for (int cc = 0; cc < m; ++cc) { A[r * m + cc] -= f * A[k * m + cc]; I[r * m + cc] -= f * I[k * m + cc]; }
      }
    }
    double* out = inv + iptr[b];
    for (int i = 0; i < m * m; ++i) out[i] = I[i];
  }
}
__global__ void __launch_bounds__(128) k_block_apply(int64_t nblk, const int64_t* __restrict__ bptr,
                                                      const int32_t* __restrict__ bdofs, const int64_t* __restrict__ iptr,
                                                      const double* __restrict__ inv, const double* __restrict__ r,
                                                      double* __restrict__ z) {
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < nblk; b += (int64_t)gridDim.x * blockDim.x) {
    const int64_t d0 = bptr[b];
    const int m = (int)(bptr[b + 1] - d0);
    double rl[FORM_MAX_BLOCK];
    for (int i = 0; i < m; ++i) rl[i] = r[bdofs[d0 + i]];
    const double* B = inv + iptr[b];
    for (int i = 0; i < m; ++i) {
      double s = 0.0;
      for (int j = 0; j < m; ++j) s += B[i * m + j] * rl[j];
      z[bdofs[d0 + i]] = s;
    }
  }
}

// ================================================================================================
// Newton helpers
// w = x - lambda y; partial sums of ||lambda y||^2 and ||w||^2
__global__ void __launch_bounds__(256) k_step(int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                                              double lambda, double* __restrict__ w, int nparts,
                                              double* __restrict__ partials) {
  __shared__ double s_red[32];
  double py = 0.0, pw = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double s = lambda * y[i], v = x[i] - s;
    w[i] = v;
    py += s * s;
    pw += v * v;
  }
  const double a = lvpp_block_sum<256>(py, s_red);
  if (threadIdx.x == 0) partials[blockIdx.x] = a;
  const double b = lvpp_block_sum<256>(pw, s_red);
  if (threadIdx.x == 0) partials[nparts + blockIdx.x] = b;
}
// partial F . Jy, ||y||^2 and max_i |y_i| / max(|x_i|, 1)  (line search bt)
__global__ void __launch_bounds__(256) k_ls_init(int64_t n, const double* __restrict__ F, const double* __restrict__ Jy,
                                                  const double* __restrict__ y, const double* __restrict__ x, int nparts,
                                                  double* __restrict__ partials) {
  __shared__ double s_red[32];
  __shared__ double s_max[32];
  double pd = 0.0, py = 0.0, mx = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    pd += F[i] * Jy[i];
    py += y[i] * y[i];
    mx = fmax(mx, fabs(y[i]) / fmax(fabs(x[i]), 1.0));
  }
  const double a = lvpp_block_sum<256>(pd, s_red);
  if (threadIdx.x == 0) partials[blockIdx.x] = a;
  const double b = lvpp_block_sum<256>(py, s_red);
  if (threadIdx.x == 0) partials[nparts + blockIdx.x] = b;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m2 = 0.0;
    for (int i = 0; i < 8; ++i) m2 = fmax(m2, s_max[i]);
    partials[2 * nparts + blockIdx.x] = m2;
  }
}
__global__ void __launch_bounds__(256) k_reduce_ls(int nparts, const double* __restrict__ partials, double* __restrict__ out) {
  __shared__ double s_red[32];
  for (int v = 0; v < 2; ++v) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 256) s += partials[(int64_t)v * nparts + i];
    const double r = lvpp_block_sum<256>(s, s_red);
    if (threadIdx.x == 0) out[v] = r;
  }
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int i = 0; i < nparts; ++i) m = fmax(m, partials[2 * (int64_t)nparts + i]);
    out[2] = m;
  }
}
// increments: |sum_a d_a phi_a(q)|^2 summed over quadrature points and components
__global__ void __launch_bounds__(128) k_increment_l2(int64_t E, int nq, int nld, int nb, int ncomp, int comp_stride,
                                                       int basis_stride, const int32_t* __restrict__ dofs,
                                                       const int32_t* __restrict__ verts, int nv, int gdim,
                                                       const double* __restrict__ coords, const double* __restrict__ w,
                                                       const double* __restrict__ tab, const double* __restrict__ x,
                                                       const double* __restrict__ x0, double* __restrict__ partials) {
  __shared__ double s_red[32];
  double part = 0.0;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < E; c += (int64_t)gridDim.x * blockDim.x) {
    int32_t vt[3];
    for (int v = 0; v < 3; ++v) vt[v] = verts[c * nv + v];
    double Jinv[2][2];
    const double adet = cell_geometry<2>(coords, vt, Jinv);
    for (int m = 0; m < ncomp; ++m) {
      double dl[6];
      for (int b = 0; b < nb; ++b) {
        const int32_t dof = dofs[c * nld + b * basis_stride + m * comp_stride];
        dl[b] = x[dof] - x0[dof];
      }
      for (int q = 0; q < nq; ++q) {
        double dq = 0.0;
        for (int b = 0; b < nb; ++b) dq += dl[b] * tab[q * nb + b];
        part += w[q] * adet * dq * dq;
      }
    }
  }
  const double r = lvpp_block_sum<128>(part, s_red);
  if (threadIdx.x == 0) partials[blockIdx.x] = r;
}
__global__ void __launch_bounds__(256) k_increment_discrete(int64_t n, const double* __restrict__ x, const double* __restrict__ x0,
                                                             double* __restrict__ partials) {
  __shared__ double s_red[32];
  double part = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double d = x[i] - x0[i];
    part += d * d;
  }
  const double r = lvpp_block_sum<256>(part, s_red);
  if (threadIdx.x == 0) partials[blockIdx.x] = r;
}

// ================================================================================================
// host side
static int form_reduce(lvpp_form_problem* h, double* partials, int nvals, double* dst_host) {
  LAUNCH(h, k_reduce_multi, nvals < 64 ? nvals : 64, 256, 0, h->npartials, nvals, partials, h->gm_h);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(dst_host, h->gm_h, sizeof(double) * nvals, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

static ElemArgs elem_args(lvpp_form_problem* h, int k) {
  const IntegralDev& I = h->itg[k];
  ElemArgs a;
  a.E = I.E; a.nq = I.nq; a.dofs = I.dofs; a.verts = I.verts; a.coords = h->coords; a.tab = I.tab;
  a.off_a = I.off_a; a.off_da = I.off_da; a.off_b = I.off_b; a.tab_len = I.tab_len;
  a.xt = h->xt; a.aux0 = h->aux0; a.aux1 = h->aux1; a.coef0 = h->coef0; a.coef1 = h->coef1;
  memcpy(a.p, h->params, sizeof(a.p));
  a.cd = I.cd; a.be = I.be;
  return a;
}
static GatherArgs gather_args(lvpp_form_problem* h) {
  GatherArgs g;
  memset(&g, 0, sizeof(g));
  g.form = h->form; g.nint = h->nint;
  for (int k = 0; k < h->nint; ++k) {
    g.cd[k] = h->itg[k].cd; g.be[k] = h->itg[k].be; g.cd_stride[k] = h->itg[k].cd_stride; g.nld[k] = h->itg[k].nld;
    g.mat_off[k] = h->itg[k].mat_off; g.vec_off[k] = h->itg[k].vec_off;
  }
  g.mat_off[h->nint] = INT64_MAX; g.vec_off[h->nint] = INT64_MAX;
  if (h->nint == 1) { g.mat_off[1] = INT64_MAX; g.vec_off[1] = INT64_MAX; }
  memcpy(g.p, h->params, sizeof(g.p));
  return g;
}

// F(x) and J(x): element kernels, then the two gathers; ||F||^2 partials are reduced into gm_h_host[0]
static int form_assemble(lvpp_form_problem* h, const double* d_x, double* d_F, double* fnorm) {
  LAUNCH(h, k_apply_bc_values, lvpp_grid(h->n, 256, 8), 256, 0, h->n, d_x, h->bc_flag, h->bc_val, h->xt);
  CK(cudaGetLastError());
  for (int k = 0; k < h->nint; ++k) {
    ElemArgs a = elem_args(h, k);
    const int grid = lvpp_grid(a.E, 128, 16);
    const size_t smem = sizeof(double) * a.tab_len;
    if (h->form == LVPP_FORM_GRADIENT) LAUNCH(h, k_elem_gradient, grid, 128, smem, a);
    else if (h->form == LVPP_FORM_MULTIPHASE) LAUNCH(h, k_elem_multiphase, grid, 128, smem, a);
    else if (k == 0) LAUNCH(h, k_elem_elasticity, grid, 128, 0, a);
    else LAUNCH(h, k_elem_contact, grid, 128, smem, a);
    CK(cudaGetLastError());
  }
  GatherArgs g = gather_args(h);
  LAUNCH(h, k_gather_matrix, lvpp_grid(h->nnz, 256, 8), 256, 0, h->nnz, g, h->mptr, h->msrc, h->nnz_row, h->indices,
         h->bc_flag, h->vals);
  CK(cudaGetLastError());
  LAUNCH(h, k_gather_vector, h->npartials, 256, 0, h->n, g, h->vptr, h->vsrc, h->bc_flag, h->bc_val, d_x, d_F,
         h->npartials, h->gm_part);
  CK(cudaGetLastError());
  h->jac_valid = true;
  h->residual_evals++;
  if (fnorm) {
    CKR(form_reduce(h, h->gm_part, 1, h->gm_h_host));
    *fnorm = sqrt(h->gm_h_host[0]);
  }
  return 0;
}

static int form_spmv(lvpp_form_problem* h, const double* v, double* y) {
  const int tpr = h->spmv_tpr;
  const int grid = lvpp_grid(h->n * tpr, 256, 8);
  h->spmv_launches++;
  switch (tpr) {
    case 4: LAUNCH(h, k_csr_spmv<4>, grid, 256, 0, h->n, h->indptr, h->indices, h->vals, v, y); break;
    case 8: LAUNCH(h, k_csr_spmv<8>, grid, 256, 0, h->n, h->indptr, h->indices, h->vals, v, y); break;
    case 16: LAUNCH(h, k_csr_spmv<16>, grid, 256, 0, h->n, h->indptr, h->indices, h->vals, v, y); break;
    default: LAUNCH(h, k_csr_spmv<32>, grid, 256, 0, h->n, h->indptr, h->indices, h->vals, v, y); break;
  }
  CK(cudaGetLastError());
  return 0;
}

static int form_precond_setup(lvpp_form_problem* h) {
  int* d_err = nullptr;
  CKR(fdalloc(h, &d_err, 1));
  LAUNCH(h, k_block_inverse, lvpp_grid(h->nblk, 64, 16), 64, 0, h->nblk, h->blk_ptr, h->blk_dofs, h->blk_inv_ptr,
         h->indptr, h->indices, h->vals, h->blk_inv, d_err);
  CK(cudaGetLastError());
  int herr = 0;
  CK(cudaMemcpyAsync(&herr, d_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CKR(fdfree(h, d_err));
  if (herr) { lvpp_set_error("block Jacobi: singular diagonal block"); return LVPP_E_INVALID; }
  return 0;
}
static int form_precond(lvpp_form_problem* h, const double* r, double* z) {
  LAUNCH(h, k_block_apply, lvpp_grid(h->nblk, 128, 8), 128, 0, h->nblk, h->blk_ptr, h->blk_dofs, h->blk_inv_ptr,
         h->blk_inv, r, z);
  CK(cudaGetLastError());
  return 0;
}

static int form_ensure_gmres(lvpp_form_problem* h, int restart) {
  if (restart <= 0) restart = 200;
  // memory bound: (restart + 1) vectors
  const int64_t cap = (int64_t)(((size_t)24 << 30) / (sizeof(double) * (size_t)h->npad));
  if (restart + 1 > cap) restart = (int)std::max<int64_t>(10, cap - 1);
  if (h->gm_V && restart <= h->gm_restart) return 0;
  if (h->gm_V) {
    CKR(fdfree(h, h->gm_V)); CKR(fdfree(h, h->gm_part));
    CKR(fdfree(h, h->gm_H)); CKR(fdfree(h, h->gm_cs)); CKR(fdfree(h, h->gm_sn)); CKR(fdfree(h, h->gm_g)); CKR(fdfree(h, h->gm_yv));
  }
  h->gm_restart = restart;
  const size_t m = (size_t)restart;
  CKR(fdalloc(h, &h->gm_V, (m + 1) * h->npad));
  CKR(fdalloc(h, &h->gm_part, (m + 4) * h->npartials));
  CKR(fdalloc(h, &h->gm_H, (m + 1) * m));
  CKR(fdalloc(h, &h->gm_cs, m));
  CKR(fdalloc(h, &h->gm_sn, m));
  CKR(fdalloc(h, &h->gm_g, m + 1));
  CKR(fdalloc(h, &h->gm_yv, m));
  if (!h->gm_state) {
    CKR(fdalloc(h, &h->gm_state, 1));
    CKR(fdalloc(h, &h->gm_red, 8));
    CK(cudaMallocHost((void**)&h->gm_state_host, sizeof(GmState) * (GM_RING + 2)));
    for (int i = 0; i < GM_RING; ++i) CK(cudaEventCreateWithFlags(&h->gm_ev[i], cudaEventDisableTiming));
  }
  return 0;
}

// right-preconditioned restarted GMRES with the block-Jacobi preconditioner; vectors are n doubles padded to npad = even,
// handled as double2.  The recurrence lives on the device exactly as in lvpp_gmres_mg (multigrid.cu): Hessenberg column,
// Givens rotations, residual estimate, re-orthogonalisation and convergence decisions are taken by one-thread kernels,
// an iteration is a fixed list of launches with no host round trip, and the host reads a copy of the state one
// iteration late.  These systems are small (1e5 - 1e6 rows): round 1's six synchronisations per iteration were most of
// an iteration's time.
static int form_gmres(lvpp_form_problem* h, const double* d_rhs, double* d_y, const lvpp_newton_opts* o, int32_t* its_out,
                      int32_t* reason_out, double* rnorm_out) {
  CKR(form_ensure_gmres(h, o->ksp_restart));
  CKR(form_precond_setup(h));
  const int m = h->gm_restart, nb = h->npartials;
  const int64_t n2 = h->npad / 2, stride2 = h->npad / 2;
  double2* Vb = (double2*)h->gm_V;
  double* gpart = h->gm_part;
  const int maxit = o->ksp_max_it > 0 ? o->ksp_max_it : 10000;
  CK(cudaEventRecord(h->ev0, h->stream));
  auto vec = [&](int k) { return (double*)(Vb + (int64_t)k * stride2); };
  GmState* st = h->gm_state;
  GmState* ring = h->gm_state_host;
  {
    GmState init;
    memset(&init, 0, sizeof(init));
    init.rtol = o->ksp_rtol; init.atol = o->ksp_atol; init.eta2 = 0.5; init.maxit = maxit; init.first = 1;
    ring[GM_RING + 1] = init;
    CK(cudaMemcpyAsync(st, &ring[GM_RING + 1], sizeof(GmState), cudaMemcpyHostToDevice, h->stream));
  }
  double* hc = h->gm_h + 8;  // dot products / coefficients of the current pass
  CK(cudaMemsetAsync(d_y, 0, sizeof(double) * h->npad, h->stream));
  int reason = 0;
  bool first = true;
  GmState fin;
  memset(&fin, 0, sizeof(fin));
  while (true) {
    if (first) {
      CK(cudaMemcpyAsync(vec(0), d_rhs, sizeof(double) * h->npad, cudaMemcpyDeviceToDevice, h->stream));
      first = false;
    } else {  // r = rhs - J y (rhs and y are padded work vectors of the handle)
      CKR(form_spmv(h, d_y, vec(0)));
      LAUNCH(h, k_axpby, nb, 256, 0, n2, -1.0, (const double2*)vec(0), 0, (double2*)vec(0));
      LAUNCH(h, k_axpby, nb, 256, 0, n2, 1.0, (const double2*)d_rhs, 1, (double2*)vec(0));
      CK(cudaGetLastError());
    }
    LAUNCH(h, k_multi_dot, nb, 256, 0, n2, Vb, stride2, 0, 1, (const double2*)vec(0), nb, gpart, 1.0);
    LAUNCH(h, k_reduce_multi, 1, 256, 0, nb, 1, gpart, h->gm_red);
    LAUNCH(h, k_gm_cycle_begin, 1, 1, 0, st, h->gm_red, h->gm_g, m);
    LAUNCH(h, k_gm_scale, nb, 256, 0, n2, st, (double2*)vec(0));
    CK(cudaGetLastError());
    int j = 0;
    bool over = false;
    for (; j < m && !over; ++j) {
      const int slot = j % GM_RING;
      CKR(form_precond(h, vec(j), h->z));
      CKR(form_spmv(h, h->z, vec(j + 1)));
      double* Hcol = h->gm_H + (size_t)j * (m + 1);
      for (int pass = 0; pass < 2; ++pass) {
        for (int k0 = 0; k0 <= j; k0 += GM_CHUNK) {
          const int nv = std::min(GM_CHUNK, j + 1 - k0);
          LAUNCH(h, k_multi_dot, nb, 256, 0, n2, Vb, stride2, k0, nv, (const double2*)vec(j + 1), nb, gpart, 1.0, st, pass);
        }
        LAUNCH(h, k_reduce_multi, (j + 1) < 64 ? (j + 1) : 64, 256, 0, nb, j + 1, gpart, hc, st, pass);
        LAUNCH(h, k_gmres_update, nb, 256, 0, n2, Vb, stride2, j + 1, hc, (double2*)vec(j + 1), nb, m + 1, gpart, 1.0, st, pass);
        LAUNCH(h, k_reduce_multi, 1, 256, 0, nb, 1, gpart + (size_t)(m + 1) * nb, h->gm_red + 1, st, pass);
        LAUNCH(h, k_gm_after_pass, 1, 1, 0, st, j, pass, hc, h->gm_red + 1, Hcol, h->gm_cs, h->gm_sn, h->gm_g);
        CK(cudaGetLastError());
      }
      LAUNCH(h, k_gm_scale, nb, 256, 0, n2, st, (double2*)vec(j + 1));
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(&ring[slot], st, sizeof(GmState), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaEventRecord(h->gm_ev[slot], h->stream));
      if (j >= 1) {  // look at iteration j - 1 while iteration j keeps the device busy
        const int ps = (j - 1) % GM_RING;
        CK(cudaEventSynchronize(h->gm_ev[ps]));
        over = ring[ps].conv != 0;
      }
    }
    CK(cudaMemcpyAsync(&ring[GM_RING], st, sizeof(GmState), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    fin = ring[GM_RING];
    const int k = fin.ncols;
    if (fin.conv && fin.reason == LVPP_KSP_DIVERGED_NANORINF) { reason = fin.reason; break; }
    if (k > 0) {
      LAUNCH(h, k_gm_backsolve, 1, 1, 0, k, m, h->gm_H, h->gm_g, h->gm_yv);
      double* comb = vec(k);
      LAUNCH(h, k_lincomb, nb, 256, 0, n2, Vb, stride2, k, h->gm_yv, (double2*)comb);
      CK(cudaGetLastError());
      CKR(form_precond(h, comb, h->z));
      LAUNCH(h, k_axpby, nb, 256, 0, n2, 1.0, (const double2*)h->z, 1, (double2*)d_y);  // y += M^-1 (V yv)
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

// ------------------------------------------------------------------------------------------------
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

// w = x - lambda y, G = F(w); returns gnorm and the two step norms
static int form_trial(lvpp_form_problem* h, const double* x, double lambda, double* gnorm, double* snorm, double* wnorm) {
  LAUNCH(h, k_step, h->npartials, 256, 0, h->n, x, h->y, lambda, h->w, h->npartials, h->gm_part);
  CK(cudaGetLastError());
  double two[2];
  CKR(form_reduce(h, h->gm_part, 2, two));
  *snorm = sqrt(two[0]);
  *wnorm = sqrt(two[1]);
  CKR(form_assemble(h, h->w, h->G, gnorm));
  return 0;
}

#define CHECK_F(h)                                                       \
  do {                                                                   \
    if (!(h)) { lvpp_set_error("null handle"); return LVPP_E_INVALID; }  \
    CK(cudaSetDevice((h)->device));                                      \
  } while (0)

extern "C" int lvpp_form_newton_solve(lvpp_form_handle h, double* d_x, const lvpp_newton_opts* o, int32_t* its,
                                      int32_t* reason, double* h_fnorm, int32_t* linear_its) {
  CHECK_F(h);
  if (!d_x || !o) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  double fnorm = 0.0;
  CKR(form_assemble(h, d_x, h->F, &fnorm));
  const double fnorm0 = fnorm, ttol = fnorm * o->snes_rtol;
  int r = snes_converged(0, 0.0, 0.0, fnorm, ttol, fnorm0, o);
  int it = 0, lin = 0;
  while (!r) {
    if (it >= o->snes_max_it) { r = LVPP_SNES_DIVERGED_MAX_IT; break; }
    int32_t kits = 0, kreason = 0;
    CKR(form_gmres(h, h->F, h->y, o, &kits, &kreason, nullptr));
    lin += kits;
    if (kreason < 0) { r = LVPP_SNES_DIVERGED_LINEAR_SOLVE; break; }
    double gnorm = 0.0, snorm = 0.0, wnorm = 0.0;
    if (o->snes_linesearch == LVPP_LINESEARCH_BT) {
      // PETSc SNESLineSearchApply_BT, cubic (restated in oracle/snes.py:linesearch_bt)
      const double lsalpha = 1e-4, maxstep = 1e8, steptol = 1e-12;
      CKR(form_spmv(h, h->y, h->Jy));  // the Jacobian at x is still in place
      LAUNCH(h, k_ls_init, h->npartials, 256, 0, h->n, h->F, h->Jy, h->y, d_x, h->npartials, h->gm_part);
      LAUNCH(h, k_reduce_ls, 1, 256, 0, h->npartials, h->gm_part, h->gm_h);
      CK(cudaGetLastError());
      double three[3];
      CK(cudaMemcpyAsync(three, h->gm_h, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      double initslope = three[0], ynorm = sqrt(three[1]), rellength = three[2];
      double scale = 1.0;
      if (ynorm > maxstep) { scale = maxstep / ynorm; initslope *= scale; rellength *= scale; }
      if (initslope > 0.0) initslope = -initslope;
      if (initslope == 0.0) initslope = -1.0;
      const double minlambda = steptol / rellength;
      const double f = fnorm * fnorm;
      double lam = 1.0;
      bool ok = false;
      if (ynorm == 0.0) {  // nothing to search along: keep x
        LAUNCH(h, k_step, h->npartials, 256, 0, h->n, d_x, h->y, 0.0, h->w, h->npartials, h->gm_part);
        CK(cudaGetLastError());
        double two[2];
        CKR(form_reduce(h, h->gm_part, 2, two));
        wnorm = sqrt(two[1]);
        CK(cudaMemcpyAsync(h->G, h->F, sizeof(double) * h->n, cudaMemcpyDeviceToDevice, h->stream));
        gnorm = fnorm; snorm = 0.0; ok = true;
      } else {
        CKR(form_trial(h, d_x, lam * scale, &gnorm, &snorm, &wnorm));
        double g = gnorm * gnorm;
        if (!std::isfinite(gnorm)) { r = LVPP_SNES_DIVERGED_LINE_SEARCH; break; }
        if (0.5 * g <= 0.5 * f + lam * lsalpha * initslope) ok = true;
        else {
          double lamtemp = -initslope / (g - f - 2.0 * lam * initslope);
          double lamprev = lam, gprev = g;
          if (lamtemp > 0.5 * lam) lamtemp = 0.5 * lam;
          lam = lamtemp <= 0.1 * lam ? 0.1 * lam : lamtemp;
          for (int k = 0; k < 40; ++k) {
            if (lam <= minlambda) break;
            CKR(form_trial(h, d_x, lam * scale, &gnorm, &snorm, &wnorm));
            g = gnorm * gnorm;
            if (0.5 * g <= 0.5 * f + lam * lsalpha * initslope) { ok = true; break; }
            const double t1 = 0.5 * (g - f) - lam * initslope, t2 = 0.5 * (gprev - f) - lamprev * initslope;
            const double a = (t1 / (lam * lam) - t2 / (lamprev * lamprev)) / (lam - lamprev);
            const double b = (-lamprev * t1 / (lam * lam) + lam * t2 / (lamprev * lamprev)) / (lam - lamprev);
            double dd = b * b - 3.0 * a * initslope;
            if (dd < 0.0) dd = 0.0;
            lamtemp = a == 0.0 ? -initslope / (2.0 * b) : (-b + sqrt(dd)) / (3.0 * a);
            lamprev = lam; gprev = g;
            if (lamtemp > 0.5 * lam) lamtemp = 0.5 * lam;
            lam = lamtemp <= 0.1 * lam ? 0.1 * lam : lamtemp;
          }
        }
      }
      if (!ok) { r = LVPP_SNES_DIVERGED_LINE_SEARCH; break; }
    } else {
      CKR(form_trial(h, d_x, 1.0, &gnorm, &snorm, &wnorm));
    }
    CK(cudaMemcpyAsync(d_x, h->w, sizeof(double) * h->n, cudaMemcpyDeviceToDevice, h->stream));
    std::swap(h->F, h->G);
    fnorm = gnorm;
    ++it;
    h->newton_steps++;
    r = snes_converged(it, wnorm, snorm, fnorm, ttol, fnorm0, o);
  }
  CK(cudaStreamSynchronize(h->stream));
  if (its) *its = it;
  if (reason) *reason = r;
  if (h_fnorm) *h_fnorm = fnorm;
  if (linear_its) *linear_its = lin;
  return LVPP_OK;
}

extern "C" int lvpp_form_linear_solve(lvpp_form_handle h, const double* d_rhs, double* d_y, const lvpp_newton_opts* o,
                                      int32_t* its, int32_t* reason, double* h_rnorm) {
  CHECK_F(h);
  if (!d_rhs || !d_y || !o) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  if (!h->jac_valid) { lvpp_set_error("no Jacobian assembled yet"); return LVPP_E_INVALID; }
  // the solver works on padded vectors: stage the right-hand side
  CK(cudaMemcpyAsync(h->Jy, d_rhs, sizeof(double) * h->n, cudaMemcpyDeviceToDevice, h->stream));
  CKR(form_gmres(h, h->Jy, h->y, o, its, reason, h_rnorm));
  CK(cudaMemcpyAsync(d_y, h->y, sizeof(double) * h->n, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return LVPP_OK;
}

extern "C" int lvpp_form_assemble_residual(lvpp_form_handle h, const double* d_x, double* d_F, double* h_fnorm) {
  CHECK_F(h);
  if (!d_x || !d_F) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  double fn = 0.0;
  CK(cudaEventRecord(h->ev0, h->stream));
  CKR(form_assemble(h, d_x, d_F, &fn));
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->t_assembly_ms += ms;
  if (h_fnorm) *h_fnorm = fn;
  return LVPP_OK;
}

extern "C" int lvpp_form_get_jacobian_values(lvpp_form_handle h, double* d_values) {
  CHECK_F(h);
  if (!d_values) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  if (!h->jac_valid) { lvpp_set_error("no Jacobian assembled yet"); return LVPP_E_INVALID; }
  CK(cudaMemcpyAsync(d_values, h->vals, sizeof(double) * h->nnz, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return LVPP_OK;
}

extern "C" int lvpp_form_spmv(lvpp_form_handle h, const double* d_v, double* d_y) {
  CHECK_F(h);
  if (!d_v || !d_y) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  if (!h->jac_valid) { lvpp_set_error("no Jacobian assembled yet"); return LVPP_E_INVALID; }
  CKR(form_spmv(h, d_v, d_y));
  CK(cudaStreamSynchronize(h->stream));
  return LVPP_OK;
}

extern "C" int lvpp_form_set_param(lvpp_form_handle h, int32_t index, double value) {
  if (!h || index < 0 || index >= FORM_MAX_PARAMS) { lvpp_set_error("bad argument"); return LVPP_E_INVALID; }
  h->params[index] = value;
  h->jac_valid = false;
  return LVPP_OK;
}

extern "C" int lvpp_form_set_aux(lvpp_form_handle h, int32_t which, const double* d_v) {
  CHECK_F(h);
  if (!d_v || which < 0 || which > 1) { lvpp_set_error("bad argument"); return LVPP_E_INVALID; }
  CK(cudaMemcpyAsync(which == 0 ? h->aux0 : h->aux1, d_v, sizeof(double) * h->n, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return LVPP_OK;
}

__global__ void k_set_bc_values(int64_t nbc, const int64_t* __restrict__ dofs, const double* __restrict__ v,
                                uint8_t* __restrict__ flag, double* __restrict__ val) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nbc; p += (int64_t)gridDim.x * blockDim.x) {
    flag[dofs[p]] = 1;
    val[dofs[p]] = v[p];
  }
}

extern "C" int lvpp_form_set_bc_values(lvpp_form_handle h, const double* h_values) {
  CHECK_F(h);
  if (!h_values && h->num_bc > 0) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  if (h->num_bc == 0) return LVPP_OK;
  double* dv = nullptr;
  CKR(fupload(h, &dv, h_values, (size_t)h->num_bc));
  LAUNCH(h, k_set_bc_values, lvpp_grid(h->num_bc, 256, 4), 256, 0, h->num_bc, h->bc_dofs, dv, h->bc_flag, h->bc_val);
  CK(cudaGetLastError());
  CKR(fdfree(h, dv));
  return LVPP_OK;
}

extern "C" int lvpp_form_increment_sq(lvpp_form_handle h, const double* d_x, const double* d_x0, double* h_out) {
  CHECK_F(h);
  if (!d_x || !d_x0 || !h_out) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  const IntegralDev& I = h->itg[0];
  if (h->form == LVPP_FORM_GRADIENT) {
    // u - u0 in P2: local dofs 0..5, one component
    LAUNCH(h, k_increment_l2, h->npartials, 128, 0, I.E, I.nq, I.nld, 6, 1, 0, 1, I.dofs, I.verts, I.nv, h->gdim, h->coords,
           I.tab, I.tab + I.off_a, d_x, d_x0, h->gm_part);
  } else if (h->form == LVPP_FORM_MULTIPHASE) {
    // the 4 species of u: local dof v * 12 + m
    LAUNCH(h, k_increment_l2, h->npartials, 128, 0, I.E, I.nq, I.nld, 3, 4, 1, 12, I.dofs, I.verts, I.nv, h->gdim, h->coords,
           I.tab, I.tab + I.off_a, d_x, d_x0, h->gm_part);
  } else {
    // discrete l2 norm over the displacement dofs (the first gdim * num_vertices rows)
    LAUNCH(h, k_increment_discrete, h->npartials, 256, 0, (int64_t)h->gdim * h->nverts, d_x, d_x0, h->gm_part);
  }
  CK(cudaGetLastError());
  CKR(form_reduce(h, h->gm_part, 1, h_out));
  return LVPP_OK;
}

extern "C" int lvpp_form_get_stats(lvpp_form_handle h, lvpp_stats* s) {
  if (!h || !s) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  memset(s, 0, sizeof(*s));
  s->num_rows = h->n; s->local_rows = h->n; s->nnz = h->nnz; s->scalar_nnz = h->nnz; s->sell_slots = h->nnz;
  s->krylov_iterations = h->krylov_its; s->newton_steps = h->newton_steps; s->residual_evals = h->residual_evals;
  s->kernel_launches = h->launches; s->device_bytes = h->device_bytes;
  s->t_assembly_ms = h->t_assembly_ms; s->t_krylov_ms = h->t_krylov_ms;
  s->fine_op_launches = h->spmv_launches;
  return LVPP_OK;
}

extern "C" int lvpp_form_time_kernels(lvpp_form_handle h, const double* d_x, int32_t reps, double* h_ms_assembly,
                                      double* h_ms_spmv) {
  CHECK_F(h);
  if (!d_x || reps < 1) { lvpp_set_error("bad argument"); return LVPP_E_INVALID; }
  float ms = 0.f;
  CKR(form_assemble(h, d_x, h->G, nullptr));  // warm-up
  CK(cudaEventRecord(h->ev0, h->stream));
  for (int r = 0; r < reps; ++r) CKR(form_assemble(h, d_x, h->G, nullptr));
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaEventSynchronize(h->ev1));
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  if (h_ms_assembly) *h_ms_assembly = ms / reps;
  CKR(form_spmv(h, d_x, h->Jy));
  CK(cudaEventRecord(h->ev0, h->stream));
  for (int r = 0; r < reps; ++r) CKR(form_spmv(h, d_x, h->Jy));
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaEventSynchronize(h->ev1));
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  if (h_ms_spmv) *h_ms_spmv = ms / reps;
  return LVPP_OK;
}

// ------------------------------------------------------------------------------------------------
// contribution lists: sort (destination, flat source index) pairs, then segment pointers
static int build_contrib(lvpp_form_problem* h, int64_t ndst, int64_t m, int64_t* d_keys, int64_t** ptr_out, uint32_t** src_out) {
  if (m >= (int64_t)0xffffffffLL) { lvpp_set_error("too many element contributions for 32-bit source indices"); return LVPP_E_CAPACITY; }
  int64_t* skeys = nullptr;
  uint32_t *vals = nullptr, *svals = nullptr;
  CKR(fdalloc(h, &skeys, (size_t)m, false));
  CKR(fdalloc(h, &vals, (size_t)m, false));
  CKR(fdalloc(h, &svals, (size_t)m, false));
  LAUNCH(h, k_iota_u32, lvpp_grid(m, 256, 16), 256, 0, m, 0u, vals);
  CK(cudaGetLastError());
  int bits = 1;
  while (((int64_t)1 << bits) <= ndst && bits < 63) ++bits;
  size_t tb = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, d_keys, skeys, vals, svals, m, 0, bits, h->stream));
  void* tmp = nullptr;
  CKR(fdalloc(h, (char**)&tmp, tb, false));
  CK(cub::DeviceRadixSort::SortPairs(tmp, tb, d_keys, skeys, vals, svals, m, 0, bits, h->stream));
  int64_t* ptr = nullptr;
  CKR(fdalloc(h, &ptr, (size_t)ndst + 1, false));
  LAUNCH(h, k_segment_ptr, lvpp_grid(ndst + 1, 256, 16), 256, 0, ndst, m, skeys, ptr);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  CKR(fdfree(h, skeys));
  CKR(fdfree(h, vals));
  CKR(fdfree(h, tmp));
  *ptr_out = ptr;
  *src_out = svals;
  return 0;
}

extern "C" int lvpp_form_create(const lvpp_form_desc* d, lvpp_form_handle* out) {
  if (!d || !out) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  *out = nullptr;
  int ndev = lvpp_device_count();
  if (ndev < 0) return ndev;
  if (d->form < LVPP_FORM_GRADIENT || d->form > LVPP_FORM_SIGNORINI) { lvpp_set_error("unknown form %d", d->form); return LVPP_E_INVALID; }
  const int want_int = d->form == LVPP_FORM_SIGNORINI ? 2 : 1;
  if (d->num_integrals != want_int || !d->integrals) { lvpp_set_error("form %d needs %d integral(s)", d->form, want_int); return LVPP_E_INVALID; }
  if (d->num_dofs < 1 || d->num_dofs >= (int64_t)0x7fffffff || !d->indptr || !d->indices || !d->vertex_coords ||
      d->num_params < 1 || d->num_params > FORM_MAX_PARAMS || !d->params || d->num_blocks < 1 || !d->block_ptr || !d->block_dofs) {
    lvpp_set_error("invalid form descriptor");
    return LVPP_E_INVALID;
  }
  const int want_gdim = d->form == LVPP_FORM_SIGNORINI ? 3 : 2;
  if (d->gdim != want_gdim) { lvpp_set_error("form %d needs gdim %d", d->form, want_gdim); return LVPP_E_INVALID; }
  if (d->form == LVPP_FORM_GRADIENT && (!d->coef0 || !d->coef1)) { lvpp_set_error("gradient form needs coef0 (phi) and coef1 (f)"); return LVPP_E_INVALID; }
  for (int64_t b = 0; b < d->num_blocks; ++b)
    if (d->block_ptr[b + 1] - d->block_ptr[b] < 1 || d->block_ptr[b + 1] - d->block_ptr[b] > FORM_MAX_BLOCK) {
      lvpp_set_error("preconditioner block %lld has an unsupported size", (long long)b);
      return LVPP_E_INVALID;
    }
  lvpp_form_problem* h = new lvpp_form_problem();
  cudaGetDevice(&h->device);
  int rc = [&]() -> int {
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&h->ev0));
    CK(cudaEventCreate(&h->ev1));
    h->form = d->form; h->gdim = d->gdim; h->n = d->num_dofs; h->npad = d->num_dofs + (d->num_dofs & 1);
    h->nverts = d->num_vertices;
    h->nnz = d->indptr[d->num_dofs];
    memcpy(h->params, d->params, sizeof(double) * d->num_params);
    h->npartials = LVPP_NUM_SMS * 4;
    CKR(fupload(h, &h->indptr, d->indptr, (size_t)h->n + 1));
    CKR(fupload(h, &h->indices, d->indices, (size_t)h->nnz));
    CKR(fdalloc(h, &h->nnz_row, (size_t)h->nnz, false));
    CKR(fdalloc(h, &h->vals, (size_t)h->nnz));
    LAUNCH(h, k_nnz_rows, lvpp_grid(h->n, 256, 16), 256, 0, h->n, h->indptr, h->nnz_row);
    CK(cudaGetLastError());
    CKR(fupload(h, &h->coords, d->vertex_coords, (size_t)d->num_vertices * d->gdim));
    CKR(fdalloc(h, &h->bc_flag, (size_t)h->n));
    CKR(fdalloc(h, &h->bc_val, (size_t)h->n));
    h->num_bc = d->num_bc;
    if (d->num_bc > 0) {
      if (!d->bc_dofs) { lvpp_set_error("bc_dofs is null"); return LVPP_E_INVALID; }
      for (int64_t p = 0; p < d->num_bc; ++p)
        if (d->bc_dofs[p] < 0 || d->bc_dofs[p] >= h->n) { lvpp_set_error("bc dof out of range"); return LVPP_E_INVALID; }
      CKR(fupload(h, &h->bc_dofs, d->bc_dofs, (size_t)d->num_bc));
      std::vector<double> zeros;
      const double* bv = d->bc_values;
      if (!bv) { zeros.assign((size_t)d->num_bc, 0.0); bv = zeros.data(); }
      CK(cudaStreamSynchronize(h->stream));
      CKR(lvpp_form_set_bc_values(h, bv));
    }
    double** vecs[] = {&h->aux0, &h->aux1, &h->coef0, &h->coef1, &h->xt, &h->F, &h->y, &h->w, &h->G, &h->Jy, &h->z};
    for (double** v : vecs) CKR(fdalloc(h, v, (size_t)h->npad));
    if (d->coef0) CK(cudaMemcpyAsync(h->coef0, d->coef0, sizeof(double) * h->n, cudaMemcpyHostToDevice, h->stream));
    if (d->coef1) CK(cudaMemcpyAsync(h->coef1, d->coef1, sizeof(double) * h->n, cudaMemcpyHostToDevice, h->stream));
    CKR(fdalloc(h, &h->gm_h, 1024));
    CK(cudaMallocHost((void**)&h->gm_h_host, sizeof(double) * 64));
    // integrals
    h->nint = d->num_integrals;
    int64_t mat_total = 0, vec_total = 0;
    for (int k = 0; k < h->nint; ++k) {
      const lvpp_integral_desc& s = d->integrals[k];
      IntegralDev& I = h->itg[k];
      int want_nld = 0, want_nv = 0, cds = 0, na = 0, nda = 0, nbb = 0;
      if (d->form == LVPP_FORM_GRADIENT) { want_nld = 12; want_nv = 3; cds = GRAD_CD; na = 6; nda = 12; nbb = 3; }
      else if (d->form == LVPP_FORM_MULTIPHASE) { want_nld = 36; want_nv = 3; cds = MP_CD; na = 3; }
      else if (k == 0) { want_nld = 12; want_nv = 4; cds = SIG_CD0; }
      else { want_nld = 12; want_nv = 3; cds = SIG_CD1; na = 3; }
      if (s.nld != want_nld || s.nv != want_nv || s.num_entities < 1 || !s.dofs || !s.vertices || !s.to_nnz) {
        lvpp_set_error("integral %d: expected nld %d, nv %d and non-null arrays", k, want_nld, want_nv);
        return LVPP_E_INVALID;
      }
      if (na > 0 && (s.nq < 1 || s.nq > LVPP_MAX_NQ || !s.qweights || !s.tab_a || (nda && !s.dtab_a) || (nbb && !s.tab_b))) {
        lvpp_set_error("integral %d: missing quadrature tables", k);
        return LVPP_E_INVALID;
      }
      I.E = s.num_entities; I.nld = s.nld; I.nv = s.nv; I.nq = na > 0 ? s.nq : 0; I.cd_stride = cds;
      for (int64_t p = 0; p < I.E * I.nld; ++p)
        if (s.dofs[p] < 0 || s.dofs[p] >= h->n) { lvpp_set_error("integral %d: dof out of range", k); return LVPP_E_INVALID; }
      for (int64_t p = 0; p < I.E * I.nv; ++p)
        if (s.vertices[p] < 0 || s.vertices[p] >= d->num_vertices) { lvpp_set_error("integral %d: vertex out of range", k); return LVPP_E_INVALID; }
      CKR(fupload(h, &I.dofs, s.dofs, (size_t)I.E * I.nld));
      CKR(fupload(h, &I.verts, s.vertices, (size_t)I.E * I.nv));
      std::vector<double> t;
      if (na > 0) {
        t.insert(t.end(), s.qweights, s.qweights + s.nq);
        I.off_a = (int)t.size();
        t.insert(t.end(), s.tab_a, s.tab_a + (size_t)s.nq * na);
        I.off_da = (int)t.size();
        if (nda) t.insert(t.end(), s.dtab_a, s.dtab_a + (size_t)s.nq * nda);
        I.off_b = (int)t.size();
        if (nbb) t.insert(t.end(), s.tab_b, s.tab_b + (size_t)s.nq * nbb);
      }
      I.tab_len = (int)t.size();
      if (sizeof(double) * t.size() > 48 * 1024) { lvpp_set_error("quadrature tables exceed 48 KB of shared memory"); return LVPP_E_CAPACITY; }
      if (!t.empty()) CKR(fupload(h, &I.tab, t.data(), t.size()));
      CKR(fdalloc(h, &I.cd, (size_t)I.E * cds, false));
      CKR(fdalloc(h, &I.be, (size_t)I.E * I.nld, false));
      I.mat_off = mat_total;
      I.vec_off = vec_total;
      mat_total += I.E * I.nld * I.nld;
      vec_total += I.E * I.nld;
    }
    {  // contribution lists (matrix: destination = CSR position; vector: destination = row)
      int64_t* keys = nullptr;
      CKR(fdalloc(h, &keys, (size_t)mat_total, false));
      for (int k = 0; k < h->nint; ++k) {
        const IntegralDev& I = h->itg[k];
        const int64_t cnt = I.E * I.nld * I.nld;
        for (int64_t p = 0; p < cnt; ++p)
          if (d->integrals[k].to_nnz[p] < 0 || d->integrals[k].to_nnz[p] >= h->nnz) { lvpp_set_error("integral %d: to_nnz out of range", k); return LVPP_E_INVALID; }
        CK(cudaMemcpyAsync(keys + I.mat_off, d->integrals[k].to_nnz, sizeof(int64_t) * cnt, cudaMemcpyHostToDevice, h->stream));
      }
      CKR(build_contrib(h, h->nnz, mat_total, keys, &h->mptr, &h->msrc));
      CKR(fdfree(h, keys));
      CKR(fdalloc(h, &keys, (size_t)vec_total, false));
      for (int k = 0; k < h->nint; ++k) {
        const IntegralDev& I = h->itg[k];
        LAUNCH(h, k_i32_to_i64, lvpp_grid(I.E * I.nld, 256, 16), 256, 0, I.E * I.nld, I.dofs, keys + I.vec_off);
        CK(cudaGetLastError());
      }
      CKR(build_contrib(h, h->n, vec_total, keys, &h->vptr, &h->vsrc));
      CKR(fdfree(h, keys));
    }
    // preconditioner blocks
    h->nblk = d->num_blocks;
    CKR(fupload(h, &h->blk_ptr, d->block_ptr, (size_t)h->nblk + 1));
    const int64_t nbd = d->block_ptr[h->nblk];
    if (nbd != h->n) { lvpp_set_error("preconditioner blocks must partition the dofs"); return LVPP_E_INVALID; }
    for (int64_t p = 0; p < nbd; ++p)
      if (d->block_dofs[p] < 0 || d->block_dofs[p] >= h->n) { lvpp_set_error("block dof out of range"); return LVPP_E_INVALID; }
    CKR(fupload(h, &h->blk_dofs, d->block_dofs, (size_t)nbd));
    std::vector<int64_t> iptr((size_t)h->nblk + 1, 0);
    for (int64_t b = 0; b < h->nblk; ++b) {
      const int64_t m = d->block_ptr[b + 1] - d->block_ptr[b];
      iptr[b + 1] = iptr[b] + m * m;
    }
    CKR(fupload(h, &h->blk_inv_ptr, iptr.data(), iptr.size()));
    CKR(fdalloc(h, &h->blk_inv, (size_t)iptr.back()));
    const double avg = (double)h->nnz / (double)h->n;
    h->spmv_tpr = avg > 48 ? 32 : (avg > 24 ? 16 : (avg > 10 ? 8 : 4));
    CKR(form_ensure_gmres(h, 30));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
  }();
  if (rc != 0) {
    lvpp_form_destroy(h);
    return rc;
  }
  *out = h;
  return LVPP_OK;
}

extern "C" int lvpp_form_destroy(lvpp_form_handle h) {
  if (!h) return LVPP_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto& p : h->allocs) cudaFree(p.first);
  if (h->gm_h_host) cudaFreeHost(h->gm_h_host);
  if (h->gm_state_host) cudaFreeHost(h->gm_state_host);
  for (int i = 0; i < GM_RING; ++i)
    if (h->gm_ev[i]) cudaEventDestroy(h->gm_ev[i]);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return LVPP_OK;
}
