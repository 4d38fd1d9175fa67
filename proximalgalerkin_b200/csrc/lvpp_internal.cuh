// Internal definitions shared by the translation units of liblvpp_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <utility>
#include <vector>

#include "../../include/lvpp_b200.h"

#define LVPP_MAX_NQ 64
#define LVPP_MAX_ROW 255      // longest scalar row (uint8 slot offsets)
#define LVPP_SLICE 32         // sliced-ELL slice height = warp width
#define LVPP_NUM_SMS 148      // B200
#define MG_MAX_SWEEPS 8
#define LVPP_COL_BC 0x80000000u  // bit 31 of a stored column index: column node is Dirichlet (u)

void lvpp_set_error(const char* fmt, ...);

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      lvpp_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));    \
      return LVPP_E_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define CKR(call)            \
  do {                       \
    int r_ = (call);         \
    if (r_ != 0) return r_;  \
  } while (0)

// every kernel launch of this library goes through here so that launches can be counted
#define LAUNCH(h, kern, grid, block, smem, ...)                         \
  do {                                                                  \
    kern<<<(grid), (block), (smem), (h)->stream>>>(__VA_ARGS__);        \
    (h)->launches++;                                                    \
  } while (0)

// device-resident scalars of the Krylov recurrence (no host round trips inside an iteration)
struct KryScal {
  double gamma0, gamma1, eta, s0, s1, c0, c1;
  double r0;
  double delta;
  double cv1, cv0;        // Lanczos three-term coefficients
  double inv_gamma1;      // scaling of the next Lanczos vector
  double inv_gamma_w;     // scaling used by the direction update of the current iteration
  double a1, a2, a3, tau; // direction / solution update coefficients
  double tol;
  double rtol, atol;
  double red[8];          // reduction results (allreduced across ranks)
  int conv, skip, its, reason, maxit;
};

// halo description of one level (owned nodes first, ghosts last)
struct LevelHalo {
  int num_neighbors = 0;
  std::vector<int32_t> neighbor_ranks;
  std::vector<int64_t> send_ptr, recv_ptr;
  int32_t* send_nodes = nullptr;  // device
  int32_t* recv_nodes = nullptr;  // device
  double* send_buf = nullptr;     // device [2 * nsend]
  double* recv_buf = nullptr;     // device [2 * nrecv]
  // host copies used by the multigrid setup: ghost node of every recv slot, its number on the owning
  // rank and the neighbour slot it comes from
  std::vector<int32_t> recv_nodes_host, ghost_owner_local, ghost_nbr;
  // peer-memory path (comm.cu): the pack kernel stores straight into the neighbour's receive buffer over
  // NVLink and raises a sequence flag there; the unpack kernel waits for its flags.  Double-buffered by parity.
  bool p2p = false;
  void* arena = nullptr;                       // own IPC-exported allocation: flags | rbuf[0] | rbuf[1]
  unsigned long long* flags = nullptr;         // [num_neighbors] raised by the neighbours
  double* rbuf[2] = {nullptr, nullptr};        // [2 * nrecv] each
  std::vector<double*> peer_rbuf[2];           // per neighbour: where my data goes in its rbuf[parity]
  std::vector<unsigned long long*> peer_flag;  // per neighbour: its flag for me
  std::vector<void*> peer_base;                // opened IPC mappings (closed at destroy)
  int* counters = nullptr;                     // [num_neighbors] last-block detection of the pack kernels
  unsigned long long seq = 0;
};

// one level of the aggregation multigrid hierarchy; level 0 aliases the fine operator
struct MgLevel {
  int64_t V = 0, Vown = 0;       // nodes owned + ghost, owned
  int64_t nslices = 0, slots = 0, nnz = 0;
  int64_t* slice_ptr = nullptr;
  uint32_t* col = nullptr;
  int32_t* rowlen = nullptr;
  uint8_t* diag_k = nullptr;
  double *K = nullptr, *M = nullptr, *D = nullptr;
  uint4* P = nullptr;            // packed single-precision copy {col, alpha K, M, D} read by the cycle (block_op.cuh)
  // bf16 copy (the default; LVPP_MG_PACK=fp32 for the records above): one record per PAIR of slots, 20 bytes (block_op.cuh: k_packed2_op)
  uint4* P2 = nullptr;           // {col0, col1, bf16(alpha K0) | bf16(M0), bf16(alpha K1) | bf16(M1)}
  uint32_t* Pd = nullptr;        // bf16(D0) | bf16(D1)
  float4* binv32 = nullptr;      // [Vown] single-precision copy of binv for k_packed2_op
  uint8_t* bc_flag = nullptr;    // [V] u dof of the node is inactive (Dirichlet / all-Dirichlet aggregate)
  int32_t* box = nullptr;        // [Vown * 3] integer box coordinates used by the coordinate aggregation
  // transfer to the next coarser level
  int32_t* agg = nullptr;        // [V] coarse node of every node
  int64_t* agg_ptr = nullptr;    // [Vc_own + 1]
  int32_t* agg_members = nullptr;// [Vown] owned nodes grouped by coarse node
  int64_t* gal_ptr = nullptr;    // [nnz_c + 1] segments of gal_src per coarse entry (CSR order)
  uint32_t* gal_src = nullptr;   // [slots] fine slots grouped by coarse entry, ascending inside a group
  int64_t* gal_dst = nullptr;    // [nnz_c] SELL slot of the coarse entry
  // smoother / work vectors [2V]
  double* binv = nullptr;        // [Vown * 4] inverse of the 2x2 node block, row-major
  double *b = nullptr, *x = nullptr, *t = nullptr;
  double* ev = nullptr;          // [2V] power-iteration vector (lambda_max of Binv J, kept between updates)
  double omega = 0.7;            // damping of the node-block Jacobi smoother on this level
  double sweep_omega[MG_MAX_SWEEPS] = {0};      // damping of post-smoothing sweep k (Chebyshev roots or omega)
  double sweep_omega_pre[MG_MAX_SWEEPS] = {0};  // same for the pre-smoothing sweeps
  double lambda = 0.0;           // last estimate of lambda_max(Binv J)
  bool ev_valid = false;         // ev holds the vector of the previous estimate (warm start)
  LevelHalo halo;
};

struct lvpp_problem {
  int device = 0;
  cudaStream_t stream = nullptr;
  int tdim = 0, nld = 0, nq = 0, nsym = 0;
  int64_t V = 0, Vown = 0, C = 0, Cown = 0;
  double f = 0.0, alpha = 1.0;
  int64_t launches = 0;
  int64_t device_bytes = 0;
  std::vector<std::pair<void*, size_t>> allocs;

  // mesh
  double* coords = nullptr;    // [V * tdim]
  int32_t* cells = nullptr;    // [C * nld]
  double* adetJ = nullptr;     // [C] |det J|
  double* tab = nullptr;       // device tables: w[nq], phi[nq*nld], dphi[nq*nld*tdim], qpts[nq*tdim]
  // incidence lists of the owned nodes (row gather map)
  int64_t* inc_ptr = nullptr;  // [Vown + 1]
  uint32_t* inc_val = nullptr; // [C * nld] cell * nld + local index, grouped by node, cell-sorted
  // incidence-ELL: entry t of owned node i lives at ie_ptr[i / 32] + 32 t + i % 32, so that a warp of the
  // row gather reads 32 consecutive element rows per step (setup.cu, assembly.cu:k_row_gather)
  int64_t* ie_ptr = nullptr;   // [nslices + 1]
  int64_t ie_slots = 0;
  uint32_t* pos = nullptr;     // [C * nld] incidence-ELL slot of (cell, local node); ~0u for ghost nodes
  uint8_t* ie_k = nullptr;     // [ie_slots * kb] SELL slot offset of every local column, kb = nld rounded up to 4
  // scalar pattern in sliced ELL
  int64_t nslices = 0, sell_slots = 0, scalar_nnz = 0;
  int32_t maxw = 0;
  int64_t* slice_ptr = nullptr; // [nslices + 1]
  int32_t* rowlen = nullptr;    // [Vown]
  int64_t* rowptr = nullptr;    // [Vown + 1] scalar CSR row pointer (export)
  uint32_t* col = nullptr;      // [sell_slots] column node | LVPP_COL_BC
  uint8_t* diag_k = nullptr;    // [Vown]
  double *K = nullptr, *M = nullptr, *D = nullptr;  // [sell_slots]
  double* De = nullptr;         // [ie_slots * nld] element-matrix rows in incidence-ELL order (scratch)
  // node data
  uint8_t* bc_flag = nullptr;   // [V]
  double* bc_val = nullptr;     // [V]
  double* bobs = nullptr;       // [Vown] int phi_obs phi_i
  double* fvec = nullptr;       // [Vown] int phi_i
  double* xk = nullptr;         // [2V] previous proximal iterate
  // work vectors [2V]
  double *F = nullptr, *y = nullptr, *va = nullptr, *vb = nullptr, *Az = nullptr, *za = nullptr,
         *zb = nullptr, *wa = nullptr, *wb = nullptr, *pinv = nullptr, *xhost_stage = nullptr;
  double* partials = nullptr;   // [npartials * 8]
  int npartials = 0;
  KryScal* scal = nullptr;      // device
  KryScal* scal_host = nullptr; // pinned
  double* red_host = nullptr;   // pinned [16]
  double* flush = nullptr;      // L2 flush scratch
  size_t flush_bytes = 0;
  bool jac_valid = false;
  bool y_is_newton_correction = false;  // h->y holds the solution of the last Krylov solve of a Newton step
  // Newton state
  double fnorm = 0.0;
  // stats
  int64_t krylov_its = 0, newton_steps = 0, residual_evals = 0;
  double t_assembly_ms = 0.0, t_krylov_ms = 0.0, last_spmv_ms = 0.0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evs0 = nullptr, evs1 = nullptr, evt0 = nullptr, evt1 = nullptr;
  double spmv_sampled_ms = 0.0;
  int64_t spmv_samples = 0;
  // one fine-level smoother sweep (k_packed_op, EPI_JACOBI) per V-cycle is bracketed by events (bench.py roofline)
  cudaEvent_t evp0 = nullptr, evp1 = nullptr;
  cudaEvent_t evp0_cur = nullptr, evp1_cur = nullptr;  // the pair of the iteration being queued (evp*_ring)
  bool smooth_sample_pending = false, smooth_sample_recorded = false;
  double smooth_sampled_ms = 0.0;
  int64_t smooth_samples = 0, packed_op_launches = 0;
  // communication
  int rank = 0, nranks = 1;
  void* nccl_comm = nullptr;
  int* p2p_err = nullptr;         // pinned, device-visible: set when a peer flag was not raised in time
  long long halo_timeout_ticks = 1LL << 34;  // clock64 ticks a halo kernel waits for a neighbour (~8 s; LVPP_HALO_TIMEOUT_S)
  LevelHalo halo;                 // fine-level halo (lvpp_obstacle_desc)
  int64_t global_rows = 0;
  // multigrid hierarchy + GMRES workspace (built lazily by lvpp_mg_setup)
  std::vector<MgLevel> levels;
  bool mg_ready = false;
  double h0 = 0.0;                // node spacing used by the coordinate aggregation
  double xmin[3] = {0, 0, 0};
  int mg_nsmooth = 2;
  // sweeps before / after the coarse-grid correction (LVPP_MG_NPRE / LVPP_MG_NPOST; LVPP_MG_NSMOOTH sets both).  V(2,3) with
  // the Chebyshev-root dampings kept for the whole solve: 21 Krylov iterations per Newton step late in the n = 215 solve
  // against 26 for V(2,2), 4.96 s against 6.05 s for the 37 Newton steps (profiles/r02_cycle_shape_scan.txt)
  int mg_npre = 2, mg_npost = 3;
  bool mg_fp32 = true;            // the cycle reads the packed single-precision copy of the operator
  // over-correction of the piecewise-constant coarse correction and relative damping of the smoother
  // (omega_l = mg_omega * 2 / (1.15 lambda_max)); tuned on the n = 215 obstacle problem (profiles/r01_mg_scan.txt)
  double mg_omega = 1.0, mg_over = 1.8;
  double mg_kscale = 1.0;         // coarse stiffness divided by this at every coarsening (multigrid.cu: build_next_level)
  double mg_margin = 1.10;        // safety factor on the power-iteration estimate of lambda_max(Binv J)
  int mg_power_its = 10;
  int mg_power_boost = 1;         // multiplier of the power iterations (10 while re-estimating after a failed solve)
  int32_t mg_best_its = 0;        // fewest Krylov iterations of a converged solve on this handle (adaptive Chebyshev ratio)
  int64_t mg_retries = 0;         // Krylov solves repeated after a re-estimate
  bool mg_bf16 = true;            // bf16 pair records (10 B / slot) instead of the single-precision ones (16 B / slot)
  bool mg_cheb_adapt = false;     // LVPP_MG_CHEB_ADAPT=1: fall back to plain damping once a solve needs 1.5x the best count (krylov.cu)
  double mg_cheb = 6.0;           // > 1: Chebyshev-root damping of the sweeps over [b / mg_cheb, b]; else plain damping
  double mg_alpha_est = -1.0;     // alpha of the last smoother eigenvalue estimate
  double* coarse_lu = nullptr;    // dense inverse of the coarsest operator (all ranks' rows) [nc * nc]
  int coarse_n = 0;               // global unknowns of the coarsest level
  int64_t coarse_off = 0;         // first global coarse node of this rank
  int32_t* coarse_gmap = nullptr; // [V coarsest] global coarse node of every local node
  double* coarse_bg = nullptr;    // [coarse_n] gathered right-hand side
  double* gm_V = nullptr;         // GMRES basis [(restart + 1) * 2V]
  int gm_restart = 0;
  // classical Gram-Schmidt is repeated when less than eta2 of ||w||^2 survives the projection (Daniel et al.)
  double gm_eta2 = 0.01;         // (0.5 = Daniel's criterion; PETSc's default never repeats; profiles/r01_mg_scan.txt)
  bool gm_weight_auto = true;     // equilibrated residual norm (weight on the psi rows), multigrid.cu; LVPP_GMRES_WEIGHT=off: Euclidean
  double gm_weight = 1.0, gm_weight_alpha = -1.0;
  // device-resident GMRES control (gmres_kernels.cuh: GmState): Hessenberg, rotations, residual estimate, decision
  struct GmState* gm_state = nullptr;       // device
  struct GmState* gm_state_host = nullptr;  // pinned [GM_RING + 2]: lagged per-iteration copies, cycle-end copy, upload staging
  double *gm_H = nullptr, *gm_cs = nullptr, *gm_sn = nullptr, *gm_g = nullptr, *gm_yv = nullptr;  // device
  double* gm_red = nullptr;       // device [8] reduced norms
  cudaEvent_t gm_ev[8] = {nullptr};                  // status copy of iteration j landed (ring)
  cudaEvent_t evs0_ring[8] = {nullptr}, evs1_ring[8] = {nullptr}, evp0_ring[8] = {nullptr}, evp1_ring[8] = {nullptr};
  bool gm_warm = true;            // LVPP_GMRES_WARM=0: every Krylov solve starts from zero
  bool gm_warm_next = false;      // set by lvpp_newton_step: y still holds the previous Newton correction of this handle
  int64_t gm_warm_used = 0;       // solves that started from the previous correction
  const int* gm_skip = nullptr;   // non-null inside a GMRES iteration: the large kernels of the cycle return at once when set
  double* gm_h = nullptr;         // device [restart + 2]
  double* gm_h_host = nullptr;    // pinned
  double* gm_part = nullptr;      // partial sums [(restart + 2) * npartials]
  int64_t vcycles = 0;
  int64_t fine_op_launches = 0;
};

// Synchronise the stream and look at the peer-memory halo's error word (comm.cu: a neighbour's flag that is not raised in
// time makes k_halo_p2p give up and scatter whatever sits in the receive buffer): every host synchronisation that may
// follow a halo exchange goes through here, so stale ghosts never reach the caller silently.  The word is cleared so
// that one failure is reported once.
static inline int lvpp_sync_check_comm(lvpp_problem* h) {
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) {
    lvpp_set_error("cudaStreamSynchronize -> %s", cudaGetErrorString(e));
    return LVPP_E_CUDA;
  }
  if (h->p2p_err && *h->p2p_err) {
    *h->p2p_err = 0;
    lvpp_set_error("peer-memory halo: a neighbour's flag was not raised in time (LVPP_HALO_TIMEOUT_S, LVPP_HALO=nccl)");
    return LVPP_E_COMM;
  }
  return 0;
}

template <class T>
int lvpp_dalloc(lvpp_problem* h, T** p, size_t n, bool zero = true) {
  size_t bytes = (n ? n : 1) * sizeof(T);
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, bytes);
  if (e != cudaSuccess) {
    lvpp_set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    return LVPP_E_CUDA;
  }
  if (zero) {
    e = cudaMemsetAsync(q, 0, bytes, h->stream);
    if (e != cudaSuccess) {
      lvpp_set_error("cudaMemset failed: %s", cudaGetErrorString(e));
      return LVPP_E_CUDA;
    }
  }
  h->allocs.push_back(std::make_pair(q, bytes));
  h->device_bytes += (int64_t)bytes;
  *p = (T*)q;
  return 0;
}
int lvpp_dfree(lvpp_problem* h, void* p);

static inline int lvpp_grid(int64_t work_items, int block, int per_sm = 8) {
  int64_t need = (work_items + block - 1) / block;
  int64_t cap = (int64_t)LVPP_NUM_SMS * per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ double lvpp_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
// deterministic block sum (fixed shuffle tree); result valid in thread 0. sm: >= 32 doubles
template <int NT>
__device__ __forceinline__ double lvpp_block_sum(double v, double* sm) {
  v = lvpp_warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = lane < (NT / 32) ? sm[lane] : 0.0;
    r = lvpp_warp_sum(r);
  }
  return r;
}
__host__ __device__ __forceinline__ int lvpp_sym(int a, int b, int nld) {
  // index of (a, b), a <= b, in the packed upper triangle
  return a * nld - (a * (a - 1)) / 2 + (b - a);
}

// ------------------------------------------------------------------------------------------------
// geometry of an affine simplex: |det J| and J^{-1} (rows = reference directions)
template <int TDIM>
__device__ __forceinline__ double cell_geometry(const double* __restrict__ coords,
                                                const int32_t* __restrict__ nodes, double (*Jinv)[TDIM]) {
  double x0[TDIM], J[TDIM][TDIM];  // J[d][k] = x_{k+1}[d] - x_0[d]
#pragma unroll
  for (int d = 0; d < TDIM; ++d) x0[d] = coords[(int64_t)nodes[0] * TDIM + d];
#pragma unroll
  for (int k = 0; k < TDIM; ++k)
#pragma unroll
    for (int d = 0; d < TDIM; ++d) J[d][k] = coords[(int64_t)nodes[k + 1] * TDIM + d] - x0[d];
  double det;
  if (TDIM == 2) {
    det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double id = 1.0 / det;
    Jinv[0][0] = J[1][1] * id;  Jinv[0][1] = -J[0][1] * id;
    Jinv[1][0] = -J[1][0] * id; Jinv[1][1] = J[0][0] * id;
  } else {
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    const double id = 1.0 / det;
    Jinv[0][0] = c00 * id;
    Jinv[1][0] = c01 * id;
    Jinv[2][0] = c02 * id;
    Jinv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    Jinv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    Jinv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    Jinv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    Jinv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    Jinv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
  }
  return fabs(det);
}

// stage the quadrature / basis tables in shared memory (all threads read the same entry: broadcast)
__device__ __forceinline__ void stage_tables(double* s_tab, const double* __restrict__ tab, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) s_tab[i] = tab[i];
  __syncthreads();
}

// ---- cross-TU entry points --------------------------------------------------------------------
int lvpp_build_pattern(lvpp_problem* h);                      // setup.cu
int lvpp_build_constant_operators(lvpp_problem* h, const lvpp_obstacle_desc* d);  // assembly.cu
int lvpp_eval_residual(lvpp_problem* h, const double* d_x, double* d_F, bool want_norm, bool keep_D = false);  // assembly.cu
int lvpp_apply_jacobian(lvpp_problem* h, const double* d_v, double* d_y, const double* inv_scale,
                        double* partials, const int* skip_flag);                    // assembly.cu
int lvpp_reduce_partials(lvpp_problem* h, int nvals, double* d_out);  // assembly.cu
int lvpp_allreduce_sum(lvpp_problem* h, double* d_buf, int n);       // comm.cu
int lvpp_halo_forward_impl(lvpp_problem* h, double* d_v);            // comm.cu (fine level)
int lvpp_halo_forward_level(lvpp_problem* h, LevelHalo& H, double* d_v);  // comm.cu
int lvpp_halo_p2p_setup(lvpp_problem* h, LevelHalo& H);                  // comm.cu (collective)
// packed int32 exchange over the lists of H: d_send [nsend] in send order -> d_recv [nrecv] in recv order
int lvpp_halo_exchange_i32(lvpp_problem* h, const LevelHalo& H, const int32_t* d_send, int32_t* d_recv);  // comm.cu
int lvpp_minres(lvpp_problem* h, const double* d_rhs, double* d_y, const lvpp_newton_opts* o,
                int32_t* its, int32_t* reason, double* rnorm);       // krylov.cu
int lvpp_build_preconditioner(lvpp_problem* h, const lvpp_newton_opts* o);  // krylov.cu
void lvpp_comm_destroy(lvpp_problem* h);
int lvpp_mg_setup(lvpp_problem* h);                                  // multigrid.cu
int lvpp_mg_update(lvpp_problem* h);                                 // multigrid.cu
int lvpp_mg_reestimate(lvpp_problem* h);                             // multigrid.cu
int lvpp_mg_vcycle(lvpp_problem* h, const double* b_in, double** z_out);  // multigrid.cu
int lvpp_gmres_mg(lvpp_problem* h, const double* d_rhs, double* d_y, const lvpp_newton_opts* o, int32_t* its,
                  int32_t* reason, double* rnorm);                   // multigrid.cu
int lvpp_solve_linear(lvpp_problem* h, const double* d_rhs, double* d_y, const lvpp_newton_opts* o, int32_t* its,
                      int32_t* reason, double* rnorm);               // krylov.cu: dispatch on pc_type                             // comm.cu
