// Problem construction: device copies of the mesh, node->cell incidence lists, the scalar sparsity
// pattern in sliced-ELL form and the atomic-free row-gather map.
//
// Replaces dolfinx.fem.petsc.create_matrix (reference call site src/lvpp/problem.py:110) and the
// dofmap / Dirichlet bookkeeping of examples/01_obstacle_problem/obstacle_pg.py:68-83.  The mixed
// (u, psi) matrix has one scalar node pattern with a 2x2 block per entry, so only the node pattern
// is built; the monolithic CSR pattern is derived from it on export (lvpp_get_csr_pattern).
#include <cub/cub.cuh>

#include "lvpp_internal.cuh"

static thread_local char g_err[512] = "";
void lvpp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* lvpp_last_error(void) { return g_err; }
extern "C" int lvpp_version(void) { return 100; }
extern "C" int lvpp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    lvpp_set_error("no CUDA device visible");
    return LVPP_E_NOGPU;
  }
  return n;
}

int lvpp_dfree(lvpp_problem* h, void* p) {
  if (!p) return 0;
  for (size_t i = 0; i < h->allocs.size(); ++i)
    if (h->allocs[i].first == p) {
      h->device_bytes -= (int64_t)h->allocs[i].second;
      h->allocs[i] = h->allocs.back();
      h->allocs.pop_back();
      break;
    }
  CK(cudaFree(p));
  return 0;
}

// ---- kernels ------------------------------------------------------------------------------------
__global__ void k_fill_inc_keys(int64_t n, int nld, const int32_t* __restrict__ cells,
                                uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x) {
    keys[p] = (uint32_t)cells[p];
    vals[p] = (uint32_t)p;  // = cell * nld + local index
  }
}

// inc_ptr[i] = first sorted position whose key >= i
__global__ void k_inc_ptr(int64_t Vown, int64_t n, const uint32_t* __restrict__ keys,
                          int64_t* __restrict__ inc_ptr) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= Vown;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if ((int64_t)keys[mid] < i) lo = mid + 1; else hi = mid;
    }
    inc_ptr[i] = lo;
  }
}

// One thread per owned node: sorted unique neighbour list from the incident cells.
// pass 0: row lengths.  pass 1: SELL columns, diagonal offsets and the gather map inc_k.
template <int NLD>
__global__ void k_row_pattern(int pass, int64_t Vown, const int64_t* __restrict__ inc_ptr,
                              const uint32_t* __restrict__ inc_val,
                              const int32_t* __restrict__ cells, int32_t* __restrict__ rowlen,
                              const int64_t* __restrict__ slice_ptr, uint32_t* __restrict__ col,
                              uint8_t* __restrict__ diag_k, const int64_t* __restrict__ ie_ptr,
                              uint32_t* __restrict__ pos, uint8_t* __restrict__ ie_k, int* __restrict__ err) {
  constexpr int KB = (NLD + 3) & ~3;
  int32_t buf[LVPP_MAX_ROW + 1];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown;
       i += (int64_t)gridDim.x * blockDim.x) {
    int len = 0;
    bool overflow = false;
    const int64_t e0 = inc_ptr[i], e1 = inc_ptr[i + 1];
    for (int64_t e = e0; e < e1; ++e) {
      const int64_t c = inc_val[e] / NLD;
      for (int b = 0; b < NLD; ++b) {
        const int32_t j = cells[c * NLD + b];
        // binary search for j
        int lo = 0, hi = len;
        while (lo < hi) {
          int mid = (lo + hi) >> 1;
          if (buf[mid] < j) lo = mid + 1; else hi = mid;
        }
        if (lo < len && buf[lo] == j) continue;
        if (len >= LVPP_MAX_ROW) { overflow = true; continue; }
        for (int t = len; t > lo; --t) buf[t] = buf[t - 1];
        buf[lo] = j;
        ++len;
      }
    }
    if (overflow) atomicExch(err, 1);
    if (e0 == e1) {  // isolated node: keep a diagonal entry so the row exists
      buf[0] = (int32_t)i;
      len = 1;
    }
    if (pass == 0) {
      rowlen[i] = len;
      continue;
    }
    const int64_t s = i >> 5;
    const int lane = (int)(i & 31);
    const int64_t base = slice_ptr[s] + lane;
    const int w = (int)((slice_ptr[s + 1] - slice_ptr[s]) >> 5);
    for (int k = 0; k < w; ++k) col[base + (int64_t)k * LVPP_SLICE] = (uint32_t)(k < len ? buf[k] : (int32_t)i);
    for (int k = 0; k < len; ++k)
      if (buf[k] == (int32_t)i) diag_k[i] = (uint8_t)k;
    const int64_t ibase = ie_ptr[s] + lane;
    for (int64_t e = e0; e < e1; ++e) {
      const int64_t c = inc_val[e] / NLD;
      const int64_t slot = ibase + (e - e0) * LVPP_SLICE;
      pos[inc_val[e]] = (uint32_t)slot;
      for (int b = 0; b < NLD; ++b) {
        const int32_t j = cells[c * NLD + b];
        int lo = 0, hi = len;
        while (lo < hi) {
          int mid = (lo + hi) >> 1;
          if (buf[mid] < j) lo = mid + 1; else hi = mid;
        }
        ie_k[slot * KB + b] = (uint8_t)lo;
      }
    }
  }
}

__global__ void k_slice_width(int64_t nslices, int64_t Vown, const int32_t* __restrict__ rowlen,
                              int64_t* __restrict__ slice_slots) {
  // one warp per slice
  const int lane = threadIdx.x & 31;
  for (int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; s < nslices;
       s += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const int64_t i = s * LVPP_SLICE + lane;
    int w = i < Vown ? rowlen[i] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
    if (lane == 0) slice_slots[s] = (int64_t)w * LVPP_SLICE;
  }
}

// incidence-ELL slice widths: 32 * max number of incident cells in the slice
__global__ void k_ie_width(int64_t nslices, int64_t Vown, const int64_t* __restrict__ inc_ptr,
                           int64_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; s < nslices;
       s += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const int64_t i = s * LVPP_SLICE + lane;
    int w = i < Vown ? (int)(inc_ptr[i + 1] - inc_ptr[i]) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
    if (lane == 0) out[s] = (int64_t)w * LVPP_SLICE;
  }
}

__global__ void k_rowlen64(int64_t Vown, const int32_t* __restrict__ rowlen, int64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= Vown;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = i < Vown ? rowlen[i] : 0;
}

__global__ void k_set_bc(int64_t nbc, const int32_t* __restrict__ nodes, const double* __restrict__ vals,
                         uint8_t* __restrict__ flag, double* __restrict__ bcval) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nbc;
       p += (int64_t)gridDim.x * blockDim.x) {
    flag[nodes[p]] = 1;
    bcval[nodes[p]] = vals ? vals[p] : 0.0;
  }
}

// new values on the Dirichlet nodes given at creation (the flags do not change)
__global__ void k_set_bc_values(int64_t nbc, const int32_t* __restrict__ nodes, const double* __restrict__ vals,
                                const uint8_t* __restrict__ flag, double* __restrict__ bcval, int* __restrict__ err) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nbc;
       p += (int64_t)gridDim.x * blockDim.x) {
    if (!flag[nodes[p]]) { *err = 1; continue; }
    bcval[nodes[p]] = vals[p];
  }
}

__global__ void k_flag_cols(int64_t nslots, const uint8_t* __restrict__ flag, uint32_t* __restrict__ col) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nslots;
       p += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t j = col[p] & ~LVPP_COL_BC;
    col[p] = flag[j] ? (j | LVPP_COL_BC) : j;
  }
}

// monolithic CSR pattern of the mixed system from the node pattern
__global__ void k_export_pattern(int64_t Vown, const int64_t* __restrict__ rowptr,
                                 const int64_t* __restrict__ slice_ptr, const int32_t* __restrict__ rowlen,
                                 const uint32_t* __restrict__ col, int64_t* __restrict__ indptr,
                                 int32_t* __restrict__ indices) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t S = rowptr[i];
    const int len = rowlen[i];
    const int64_t p0 = 4 * S, p1 = 4 * S + 2 * len;
    indptr[2 * i] = p0;
    indptr[2 * i + 1] = p1;
    if (i == Vown - 1) indptr[2 * Vown] = p1 + 2 * len;
    const int64_t base = slice_ptr[i >> 5] + (i & 31);
    for (int k = 0; k < len; ++k) {
      const int32_t j = (int32_t)(col[base + (int64_t)k * LVPP_SLICE] & ~LVPP_COL_BC);
      indices[p0 + 2 * k] = 2 * j;
      indices[p0 + 2 * k + 1] = 2 * j + 1;
      indices[p1 + 2 * k] = 2 * j;
      indices[p1 + 2 * k + 1] = 2 * j + 1;
    }
  }
}

// ---- host side ------------------------------------------------------------------------------------
template <int NLD>
static int row_pattern_pass(lvpp_problem* h, int pass, int* d_err) {
  LAUNCH(h, k_row_pattern<NLD>, lvpp_grid(h->Vown, 128, 16), 128, 0, pass, h->Vown, h->inc_ptr,
         h->inc_val, h->cells, h->rowlen, h->slice_ptr, h->col, h->diag_k, h->ie_ptr, h->pos, h->ie_k, d_err);
  CK(cudaGetLastError());
  return 0;
}
static int row_pattern_dispatch(lvpp_problem* h, int pass, int* d_err) {
  switch (h->nld) {
    case 3: return row_pattern_pass<3>(h, pass, d_err);
    case 4: return row_pattern_pass<4>(h, pass, d_err);
    case 6: return row_pattern_pass<6>(h, pass, d_err);
    case 10: return row_pattern_pass<10>(h, pass, d_err);
  }
  lvpp_set_error("unsupported nld %d", h->nld);
  return LVPP_E_INVALID;
}

int lvpp_build_pattern(lvpp_problem* h) {
  const int64_t n = h->C * h->nld;
  if (n >= (int64_t)0xffffffffLL) {
    lvpp_set_error("num_cells * nld = %lld does not fit the 32-bit incidence index", (long long)n);
    return LVPP_E_CAPACITY;
  }
  // 1. incidence lists: stable radix sort of (node, cell*nld+a) by node
  uint32_t *keys_in = nullptr, *keys_out = nullptr, *vals_in = nullptr;
  CKR(lvpp_dalloc(h, &keys_in, (size_t)n, false));
  CKR(lvpp_dalloc(h, &keys_out, (size_t)n, false));
  CKR(lvpp_dalloc(h, &vals_in, (size_t)n, false));
  CKR(lvpp_dalloc(h, &h->inc_val, (size_t)n, false));
  LAUNCH(h, k_fill_inc_keys, lvpp_grid(n, 256, 16), 256, 0, n, h->nld, h->cells, keys_in, vals_in);
  CK(cudaGetLastError());
  int end_bit = 1;
  while (((int64_t)1 << end_bit) < h->V && end_bit < 32) ++end_bit;
  size_t tmp_bytes = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, vals_in, h->inc_val, n, 0,
                                     end_bit, h->stream));
  void* tmp = nullptr;
  CKR(lvpp_dalloc(h, (char**)&tmp, tmp_bytes, false));
  CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in, h->inc_val, n, 0,
                                     end_bit, h->stream));
  CKR(lvpp_dalloc(h, &h->inc_ptr, (size_t)h->Vown + 1));
  LAUNCH(h, k_inc_ptr, lvpp_grid(h->Vown + 1, 256, 16), 256, 0, h->Vown, n, keys_out, h->inc_ptr);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  CKR(lvpp_dfree(h, tmp));
  CKR(lvpp_dfree(h, keys_in));
  CKR(lvpp_dfree(h, keys_out));
  CKR(lvpp_dfree(h, vals_in));

  // 2. row lengths
  int* d_err = nullptr;
  CKR(lvpp_dalloc(h, &d_err, 1));
  CKR(lvpp_dalloc(h, &h->rowlen, (size_t)h->Vown));
  CKR(row_pattern_dispatch(h, 0, d_err));
  int herr = 0;
  CK(cudaMemcpyAsync(&herr, d_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (herr) {
    lvpp_set_error("a scalar row has more than %d entries", LVPP_MAX_ROW);
    return LVPP_E_CAPACITY;
  }
  // 3. slice pointers (exclusive scan of 32 * slice width) and the scalar CSR row pointer
  h->nslices = (h->Vown + LVPP_SLICE - 1) / LVPP_SLICE;
  int64_t* slots = nullptr;
  int64_t* len64 = nullptr;
  CKR(lvpp_dalloc(h, &slots, (size_t)h->nslices + 1));
  CKR(lvpp_dalloc(h, &len64, (size_t)h->Vown + 1));
  CKR(lvpp_dalloc(h, &h->slice_ptr, (size_t)h->nslices + 1));
  CKR(lvpp_dalloc(h, &h->rowptr, (size_t)h->Vown + 1));
  LAUNCH(h, k_slice_width, lvpp_grid(h->nslices * 32, 256, 16), 256, 0, h->nslices, h->Vown, h->rowlen, slots);
  CK(cudaGetLastError());
  LAUNCH(h, k_rowlen64, lvpp_grid(h->Vown + 1, 256, 16), 256, 0, h->Vown, h->rowlen, len64);
  CK(cudaGetLastError());
  size_t scan_bytes = 0, scan_bytes2 = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, slots, h->slice_ptr, h->nslices + 1, h->stream));
  CK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes2, len64, h->rowptr, h->Vown + 1, h->stream));
  if (scan_bytes2 > scan_bytes) scan_bytes = scan_bytes2;
  void* stmp = nullptr;
  CKR(lvpp_dalloc(h, (char**)&stmp, scan_bytes, false));
  CK(cub::DeviceScan::ExclusiveSum(stmp, scan_bytes, slots, h->slice_ptr, h->nslices + 1, h->stream));
  CK(cub::DeviceScan::ExclusiveSum(stmp, scan_bytes, len64, h->rowptr, h->Vown + 1, h->stream));
  CK(cudaMemcpyAsync(&h->sell_slots, h->slice_ptr + h->nslices, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(&h->scalar_nnz, h->rowptr + h->Vown, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CKR(lvpp_dfree(h, stmp));
  CKR(lvpp_dfree(h, slots));
  CKR(lvpp_dfree(h, len64));
  // widest slice
  {
    std::vector<int64_t> sp((size_t)h->nslices + 1);
    CK(cudaMemcpy(sp.data(), h->slice_ptr, sp.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
    int64_t mw = 0;
    for (int64_t s = 0; s < h->nslices; ++s) mw = std::max(mw, (sp[s + 1] - sp[s]) / LVPP_SLICE);
    h->maxw = (int32_t)mw;
  }
  // 4. columns, diagonal offsets, gather map
  CKR(lvpp_dalloc(h, &h->col, (size_t)h->sell_slots));
  CKR(lvpp_dalloc(h, &h->diag_k, (size_t)h->Vown));
  {  // incidence-ELL slice pointers, then the (cell, local node) -> slot map and the slot-offset bytes
    int64_t* iw = nullptr;
    CKR(lvpp_dalloc(h, &iw, (size_t)h->nslices + 1));
    CKR(lvpp_dalloc(h, &h->ie_ptr, (size_t)h->nslices + 1));
    LAUNCH(h, k_ie_width, lvpp_grid(h->nslices * 32, 256, 16), 256, 0, h->nslices, h->Vown, h->inc_ptr, iw);
    CK(cudaGetLastError());
    size_t sb = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, sb, iw, h->ie_ptr, h->nslices + 1, h->stream));
    void* st = nullptr;
    CKR(lvpp_dalloc(h, (char**)&st, sb, false));
    CK(cub::DeviceScan::ExclusiveSum(st, sb, iw, h->ie_ptr, h->nslices + 1, h->stream));
    CK(cudaMemcpyAsync(&h->ie_slots, h->ie_ptr + h->nslices, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CKR(lvpp_dfree(h, st));
    CKR(lvpp_dfree(h, iw));
    if (h->ie_slots >= (int64_t)0xffffffffLL) {
      lvpp_set_error("incidence-ELL has %lld slots, more than the 32-bit map holds", (long long)h->ie_slots);
      return LVPP_E_CAPACITY;
    }
    const int kb = (h->nld + 3) & ~3;
    CKR(lvpp_dalloc(h, &h->pos, (size_t)n, false));
    CK(cudaMemsetAsync(h->pos, 0xff, sizeof(uint32_t) * n, h->stream));  // ghost nodes: no slot
    CKR(lvpp_dalloc(h, &h->ie_k, (size_t)h->ie_slots * kb));
  }
  CKR(row_pattern_dispatch(h, 1, d_err));
  // 5. Dirichlet flag on column indices
  LAUNCH(h, k_flag_cols, lvpp_grid(h->sell_slots, 256, 16), 256, 0, h->sell_slots, h->bc_flag, h->col);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  CKR(lvpp_dfree(h, d_err));
  return 0;
}

extern "C" int lvpp_create(const lvpp_obstacle_desc* d, lvpp_handle* out) {
  if (!d || !out) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  *out = nullptr;
  int ndev = lvpp_device_count();
  if (ndev < 0) return ndev;
  if (!((d->tdim == 2 && (d->nld == 3 || d->nld == 6)) || (d->tdim == 3 && (d->nld == 4 || d->nld == 10)))) {
    lvpp_set_error("unsupported element: tdim %d nld %d", d->tdim, d->nld);
    return LVPP_E_INVALID;
  }
  if (d->nq < 1 || d->nq > LVPP_MAX_NQ) { lvpp_set_error("nq %d out of range", d->nq); return LVPP_E_INVALID; }
  if (d->num_owned < 1 || d->num_owned > d->num_nodes || d->num_cells < 1 ||
      d->num_owned_cells > d->num_cells || d->num_nodes >= (int64_t)0x7fffffff) {
    lvpp_set_error("inconsistent sizes");
    return LVPP_E_INVALID;
  }
  if (!d->node_coords || !d->cell_nodes || !d->qweights || !d->phi_tab || !d->dphi_tab || !d->qpoints) {
    lvpp_set_error("null array in descriptor");
    return LVPP_E_INVALID;
  }
  if (d->obstacle_kind == LVPP_OBSTACLE_ARRAY && !d->phi_obs_q) {
    lvpp_set_error("LVPP_OBSTACLE_ARRAY needs phi_obs_q");
    return LVPP_E_INVALID;
  }
  lvpp_problem* h = new lvpp_problem();
  cudaGetDevice(&h->device);
  int rc = [&]() -> int {
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&h->ev0));
    CK(cudaEventCreate(&h->ev1));
    CK(cudaEventCreate(&h->evs0));
    CK(cudaEventCreate(&h->evs1));
    CK(cudaEventCreate(&h->evp0));
    CK(cudaEventCreate(&h->evp1));
    CK(cudaEventCreate(&h->evt0));
    CK(cudaEventCreate(&h->evt1));
    h->tdim = d->tdim; h->nld = d->nld; h->nq = d->nq;
    h->nsym = d->nld * (d->nld + 1) / 2;
    h->V = d->num_nodes; h->Vown = d->num_owned; h->C = d->num_cells; h->Cown = d->num_owned_cells;
    h->f = d->f;
    h->global_rows = 2 * h->Vown;
    CKR(lvpp_dalloc(h, &h->coords, (size_t)h->V * h->tdim, false));
    CKR(lvpp_dalloc(h, &h->cells, (size_t)h->C * h->nld, false));
    CK(cudaMemcpyAsync(h->coords, d->node_coords, sizeof(double) * h->V * h->tdim, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->cells, d->cell_nodes, sizeof(int32_t) * h->C * h->nld, cudaMemcpyHostToDevice, h->stream));
    // tables: w | phi | dphi | qpts
    {
      const int nq = h->nq, nld = h->nld, td = h->tdim;
      std::vector<double> t((size_t)nq * (1 + nld + nld * td + td));
      double* p = t.data();
      for (int i = 0; i < nq; ++i) *p++ = d->qweights[i];
      for (int i = 0; i < nq * nld; ++i) *p++ = d->phi_tab[i];
      for (int i = 0; i < nq * nld * td; ++i) *p++ = d->dphi_tab[i];
      for (int i = 0; i < nq * td; ++i) *p++ = d->qpoints[i];
      CKR(lvpp_dalloc(h, &h->tab, t.size(), false));
      CK(cudaMemcpy(h->tab, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    // Dirichlet data
    CKR(lvpp_dalloc(h, &h->bc_flag, (size_t)h->V));
    CKR(lvpp_dalloc(h, &h->bc_val, (size_t)h->V));
    if (d->num_bc > 0) {
      if (!d->bc_nodes) { lvpp_set_error("bc_nodes is null"); return LVPP_E_INVALID; }
      for (int64_t p = 0; p < d->num_bc; ++p)
        if (d->bc_nodes[p] < 0 || d->bc_nodes[p] >= h->V) { lvpp_set_error("bc node out of range"); return LVPP_E_INVALID; }
      int32_t* dn = nullptr; double* dv = nullptr;
      CKR(lvpp_dalloc(h, &dn, (size_t)d->num_bc, false));
      CK(cudaMemcpyAsync(dn, d->bc_nodes, sizeof(int32_t) * d->num_bc, cudaMemcpyHostToDevice, h->stream));
      if (d->bc_values) {
        CKR(lvpp_dalloc(h, &dv, (size_t)d->num_bc, false));
        CK(cudaMemcpyAsync(dv, d->bc_values, sizeof(double) * d->num_bc, cudaMemcpyHostToDevice, h->stream));
      }
      LAUNCH(h, k_set_bc, lvpp_grid(d->num_bc, 256), 256, 0, d->num_bc, dn, dv, h->bc_flag, h->bc_val);
      CK(cudaGetLastError());
      CK(cudaStreamSynchronize(h->stream));
      CKR(lvpp_dfree(h, dn));
      if (dv) CKR(lvpp_dfree(h, dv));
    }
    // halo lists
    LevelHalo& H = h->halo;
    H.num_neighbors = d->num_neighbors;
    if (d->num_neighbors > 0) {
      if (!d->neighbor_ranks || !d->send_ptr || !d->recv_ptr || !d->send_nodes || !d->recv_nodes) {
        lvpp_set_error("null halo array"); return LVPP_E_INVALID;
      }
      H.neighbor_ranks.assign(d->neighbor_ranks, d->neighbor_ranks + d->num_neighbors);
      H.send_ptr.assign(d->send_ptr, d->send_ptr + d->num_neighbors + 1);
      H.recv_ptr.assign(d->recv_ptr, d->recv_ptr + d->num_neighbors + 1);
      const int64_t ns = H.send_ptr.back(), nr = H.recv_ptr.back();
      for (int64_t p = 0; p < ns; ++p)
        if (d->send_nodes[p] < 0 || d->send_nodes[p] >= h->Vown) { lvpp_set_error("send node is not owned"); return LVPP_E_INVALID; }
      for (int64_t p = 0; p < nr; ++p)
        if (d->recv_nodes[p] < h->Vown || d->recv_nodes[p] >= h->V) { lvpp_set_error("recv node is not a ghost"); return LVPP_E_INVALID; }
      H.recv_nodes_host.assign(d->recv_nodes, d->recv_nodes + nr);
      CKR(lvpp_dalloc(h, &H.send_nodes, (size_t)ns, false));
      CKR(lvpp_dalloc(h, &H.recv_nodes, (size_t)nr, false));
      CK(cudaMemcpy(H.send_nodes, d->send_nodes, sizeof(int32_t) * ns, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(H.recv_nodes, d->recv_nodes, sizeof(int32_t) * nr, cudaMemcpyHostToDevice));
      CKR(lvpp_dalloc(h, &H.send_buf, (size_t)2 * ns));
      CKR(lvpp_dalloc(h, &H.recv_buf, (size_t)2 * nr));
    }
    CK(cudaStreamSynchronize(h->stream));
    {  // node spacing and origin for the coordinate aggregation of the multigrid hierarchy
      const int td = h->tdim, nv = td + 1;
      for (int dd = 0; dd < 3; ++dd) h->xmin[dd] = 0.0;
      for (int dd = 0; dd < td; ++dd) {
        double mn = d->node_coords[dd];
        for (int64_t i = 1; i < h->Vown; ++i) mn = std::min(mn, d->node_coords[i * td + dd]);
        h->xmin[dd] = mn;
      }
      const int64_t nsample = std::min<int64_t>(h->C, 200000);
      const int64_t step = std::max<int64_t>(1, h->C / nsample);
      std::vector<double> mins;
      mins.reserve((size_t)nsample + 1);
      for (int64_t c = 0; c < h->C; c += step) {
        double best = 1e300;
        for (int a = 0; a < nv; ++a)
          for (int b = a + 1; b < nv; ++b) {
            double s2 = 0.0;
            for (int dd = 0; dd < td; ++dd) {
              const double t = d->node_coords[(int64_t)d->cell_nodes[c * h->nld + a] * td + dd] -
                               d->node_coords[(int64_t)d->cell_nodes[c * h->nld + b] * td + dd];
              s2 += t * t;
            }
            best = std::min(best, s2);
          }
        mins.push_back(std::sqrt(best));
      }
      std::nth_element(mins.begin(), mins.begin() + mins.size() / 2, mins.end());
      h->h0 = mins[mins.size() / 2] / (h->nld > nv ? 2.0 : 1.0);  // P2: edge midpoints halve the spacing
      if (!(h->h0 > 0.0)) { lvpp_set_error("degenerate mesh (zero edge length)"); return LVPP_E_INVALID; }
    }
    CKR(lvpp_build_pattern(h));
    // operator storage
    CKR(lvpp_dalloc(h, &h->K, (size_t)h->sell_slots));
    CKR(lvpp_dalloc(h, &h->M, (size_t)h->sell_slots));
    CKR(lvpp_dalloc(h, &h->D, (size_t)h->sell_slots));
    CKR(lvpp_dalloc(h, &h->De, (size_t)h->ie_slots * h->nld, false));
    CKR(lvpp_dalloc(h, &h->adetJ, (size_t)h->C, false));
    CKR(lvpp_dalloc(h, &h->bobs, (size_t)h->Vown));
    CKR(lvpp_dalloc(h, &h->fvec, (size_t)h->Vown));
    const size_t n2 = (size_t)2 * h->V;
    CKR(lvpp_dalloc(h, &h->xk, n2));
    double** vecs[] = {&h->F, &h->y, &h->va, &h->vb, &h->Az, &h->za, &h->zb, &h->wa, &h->wb, &h->pinv, &h->xhost_stage};
    for (double** v : vecs) CKR(lvpp_dalloc(h, v, n2));
    h->npartials = LVPP_NUM_SMS * 6;  // k_block_op<0>: 40 registers x 256 threads -> 6 CTAs per SM
    CKR(lvpp_dalloc(h, &h->partials, (size_t)h->npartials * 8));
    CKR(lvpp_dalloc(h, &h->scal, 1));
    CK(cudaMallocHost((void**)&h->scal_host, sizeof(KryScal)));
    CK(cudaMallocHost((void**)&h->red_host, 16 * sizeof(double)));
    CKR(lvpp_build_constant_operators(h, d));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
  }();
  if (rc != 0) {
    lvpp_destroy(h);
    return rc;
  }
  *out = h;
  return LVPP_OK;
}

extern "C" int lvpp_destroy(lvpp_handle h) {
  if (!h) return LVPP_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  lvpp_comm_destroy(h);
  for (auto& p : h->allocs) cudaFree(p.first);
  if (h->flush) cudaFree(h->flush);
  if (h->scal_host) cudaFreeHost(h->scal_host);
  if (h->red_host) cudaFreeHost(h->red_host);
  if (h->gm_h_host) cudaFreeHost(h->gm_h_host);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->evs0) cudaEventDestroy(h->evs0);
  if (h->evs1) cudaEventDestroy(h->evs1);
  if (h->evp0) cudaEventDestroy(h->evp0);
  if (h->evp1) cudaEventDestroy(h->evp1);
  for (int i = 0; i < 8; ++i) {
    if (h->gm_ev[i]) cudaEventDestroy(h->gm_ev[i]);
    if (h->evs0_ring[i]) cudaEventDestroy(h->evs0_ring[i]);
    if (h->evs1_ring[i]) cudaEventDestroy(h->evs1_ring[i]);
    if (h->evp0_ring[i]) cudaEventDestroy(h->evp0_ring[i]);
    if (h->evp1_ring[i]) cudaEventDestroy(h->evp1_ring[i]);
  }
  if (h->gm_state_host) cudaFreeHost(h->gm_state_host);
  if (h->evt0) cudaEventDestroy(h->evt0);
  if (h->evt1) cudaEventDestroy(h->evt1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return LVPP_OK;
}

extern "C" int lvpp_get_stats(lvpp_handle h, lvpp_stats* s) {
  if (!h || !s) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  s->num_rows = h->global_rows;
  s->local_rows = 2 * h->Vown;
  s->nnz = 4 * h->scalar_nnz;
  s->scalar_nnz = h->scalar_nnz;
  s->sell_slots = h->sell_slots;
  s->krylov_iterations = h->krylov_its;
  s->newton_steps = h->newton_steps;
  s->residual_evals = h->residual_evals;
  s->kernel_launches = h->launches;
  s->device_bytes = h->device_bytes;
  s->t_assembly_ms = h->t_assembly_ms;
  s->t_krylov_ms = h->t_krylov_ms;
  s->last_spmv_ms = h->last_spmv_ms;
  s->spmv_sampled_ms = h->spmv_sampled_ms;
  s->spmv_samples = h->spmv_samples;
  s->fine_op_launches = h->fine_op_launches;
  s->smooth_sampled_ms = h->smooth_sampled_ms;
  s->smooth_samples = h->smooth_samples;
  s->packed_op_launches = h->packed_op_launches;
  s->vcycles = h->vcycles;
  s->mg_levels = (int32_t)h->levels.size();
  s->reserved0 = 0;
  return LVPP_OK;
}

extern "C" int lvpp_timer_start(lvpp_handle h) {
  if (!h) { lvpp_set_error("null handle"); return LVPP_E_INVALID; }
  CK(cudaSetDevice(h->device));
  CK(cudaEventRecord(h->evt0, h->stream));
  return LVPP_OK;
}
extern "C" int lvpp_timer_stop(lvpp_handle h, double* h_ms) {
  if (!h || !h_ms) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  CK(cudaSetDevice(h->device));
  CK(cudaEventRecord(h->evt1, h->stream));
  CK(cudaEventSynchronize(h->evt1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, h->evt0, h->evt1));
  *h_ms = ms;
  return LVPP_OK;
}

extern "C" int lvpp_get_csr_pattern(lvpp_handle h, int64_t* h_indptr, int32_t* h_indices) {
  if (!h || !h_indptr || !h_indices) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  CK(cudaSetDevice(h->device));
  int64_t* dptr = nullptr; int32_t* dind = nullptr;
  const size_t nnz = (size_t)4 * h->scalar_nnz;
  CKR(lvpp_dalloc(h, &dptr, (size_t)2 * h->Vown + 1));
  CKR(lvpp_dalloc(h, &dind, nnz, false));
  LAUNCH(h, k_export_pattern, lvpp_grid(h->Vown, 128, 16), 128, 0, h->Vown, h->rowptr, h->slice_ptr,
         h->rowlen, h->col, dptr, dind);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h_indptr, dptr, sizeof(int64_t) * (2 * h->Vown + 1), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(h_indices, dind, sizeof(int32_t) * nnz, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CKR(lvpp_dfree(h, dptr));
  CKR(lvpp_dfree(h, dind));
  return LVPP_OK;
}


// u_bc.x.array[...] = ... between solves (a dolfinx DirichletBC reads its Function / Constant at assembly time;
// examples/02_signorini/signorini_dolfinx.py:322 does this, the obstacle driver keeps g = 0): new values g on the
// Dirichlet nodes given to lvpp_create.  The set of Dirichlet nodes itself is part of the pattern and cannot change.
extern "C" int lvpp_set_bc_values(lvpp_handle h, int64_t num_bc, const int32_t* h_bc_nodes, const double* h_bc_values) {
  if (!h) { lvpp_set_error("null handle"); return LVPP_E_INVALID; }
  if (num_bc < 0 || (num_bc > 0 && (!h_bc_nodes || !h_bc_values))) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  if (num_bc == 0) return LVPP_OK;
  CK(cudaSetDevice(h->device));
  for (int64_t p = 0; p < num_bc; ++p)
    if (h_bc_nodes[p] < 0 || h_bc_nodes[p] >= h->V) { lvpp_set_error("bc node out of range"); return LVPP_E_INVALID; }
  int32_t* dn = nullptr; double* dv = nullptr; int* derr = nullptr;
  CKR(lvpp_dalloc(h, &dn, (size_t)num_bc, false));
  CKR(lvpp_dalloc(h, &dv, (size_t)num_bc, false));
  CKR(lvpp_dalloc(h, &derr, 1));
  CK(cudaMemcpyAsync(dn, h_bc_nodes, sizeof(int32_t) * num_bc, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(dv, h_bc_values, sizeof(double) * num_bc, cudaMemcpyHostToDevice, h->stream));
  LAUNCH(h, k_set_bc_values, lvpp_grid(num_bc, 256), 256, 0, num_bc, dn, dv, h->bc_flag, h->bc_val, derr);
  CK(cudaGetLastError());
  int herr = 0;
  CK(cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CKR(lvpp_dfree(h, dn));
  CKR(lvpp_dfree(h, dv));
  CKR(lvpp_dfree(h, derr));
  if (herr) { lvpp_set_error("lvpp_set_bc_values: a node that is not a Dirichlet node of this handle"); return LVPP_E_INVALID; }
  h->jac_valid = false;  // the residual state of lvpp_newton_begin belongs to the old data
  return LVPP_OK;
}
