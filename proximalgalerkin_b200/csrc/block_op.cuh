// The 2x2-block operator on the sliced-ELL node pattern; one thread per owned node (= two rows).
// Shared by assembly.cu (J*v, residual), krylov.cu (MINRES) and multigrid.cu (smoother, residual on
// every level of the hierarchy: coarse levels have exactly the fine level's structure).
//
//   MODE 0 (J*v):  y_u = alpha K v_u + M v_psi,  y_psi = M v_u - D v_psi  with Dirichlet rows/columns
//                  of u replaced by the identity (assemble_matrix with bcs, src/lvpp/problem.py:76),
//                  optional input scaling *inv_scale and fused partial sum of (scaled v) . y.
//                  Epilogue `epi`:  EPI_NONE   y = J v
//                                   EPI_RESID  y = b - J v
//                                   EPI_JACOBI y = v + omega * Binv (b - J v)   (damped node-block Jacobi)
//                  (the multigrid cycle applies its own packed single-precision copy: k_packed_op below)
//   MODE 1 (F):    residual of obstacle_pg.py:116-124 with apply_lifting(x0 = x, scale -1) and
//                  set_bc(x, -1) (src/lvpp/problem.py:59-67): the linear part is evaluated at x with
//                  its Dirichlet entries replaced by g, Dirichlet rows are x - g; fused partial ||F||^2.
#pragma once
#include "lvpp_internal.cuh"

enum { EPI_NONE = 0, EPI_RESID = 1, EPI_JACOBI = 2 };

struct OpArgs {
  int64_t Vown;
  const int64_t* slice_ptr;
  const uint32_t* col;
  const double *K, *M, *D;
  const uint8_t* bc_flag;
  const double* bc_val;
  double alpha;
  const double2* v;    // MODE 0: input; MODE 1: x
  const double2* xk;   // MODE 1
  const double *bobs, *fvec;
  double f;
  const double* inv_scale;
  const int* skip_flag;  // MODE 0: device flag, non-zero = Krylov solve already converged, do nothing
  double2* y;
  double* partials;    // [gridDim] or null
  // epilogue (MODE 0)
  int epi;
  const double2* b;
  const double* binv;  // [Vown * 4]
  double omega;
};

template <int MODE>
__global__ void __launch_bounds__(256) k_block_op(OpArgs p) {
  __shared__ double s_red[32];
  double part = 0.0;
  if (MODE == 0 && p.skip_flag && *p.skip_flag) return;
  const double sc = (MODE == 0 && p.inv_scale) ? *p.inv_scale : 1.0;
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x; i0 < p.Vown; i0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i0 + threadIdx.x;
    if (i < p.Vown) {
      const int64_t s = i >> 5;
      const int64_t b0 = p.slice_ptr[s];
      const int w = (int)((p.slice_ptr[s + 1] - b0) >> 5);
      const int64_t base = b0 + (i & 31);
      double au = 0.0, ap = 0.0;
#pragma unroll 4
      for (int k = 0; k < w; ++k) {
        const int64_t idx = base + (int64_t)k * LVPP_SLICE;
        const uint32_t c = p.col[idx];
        const double kv = p.K[idx], mv = p.M[idx], dv = p.D[idx];
        const uint32_t j = c & ~LVPP_COL_BC;
        double2 vj = __ldg(&p.v[j]);
        if (MODE == 0) {
          if (c & LVPP_COL_BC) vj.x = 0.0;
          au += p.alpha * kv * vj.x + mv * vj.y;
          ap += mv * vj.x - dv * vj.y;
        } else {
          if (c & LVPP_COL_BC) vj.x = p.bc_val[j];
          const double pk = __ldg(&p.xk[j]).y;
          au += p.alpha * kv * vj.x + mv * (vj.y - pk);
          ap += mv * vj.x - dv;
        }
      }
      const double2 vi = p.v[i];
      const bool isbc = p.bc_flag[i] != 0;
      double2 out;
      if (MODE == 0) {
        out.x = isbc ? vi.x * sc : au * sc;
        out.y = ap * sc;
        part += (out.x * vi.x + out.y * vi.y) * sc;
        if (p.epi != EPI_NONE) {
          const double2 bi = p.b[i];
          const double ru = bi.x - out.x, rp = bi.y - out.y;
          if (p.epi == EPI_RESID) {
            out.x = ru;
            out.y = rp;
          } else {
            const double* B = p.binv + 4 * i;
            out.x = vi.x + (isbc ? 1.0 : p.omega) * (B[0] * ru + B[1] * rp);  // Dirichlet rows exactly
            out.y = vi.y + p.omega * (B[2] * ru + B[3] * rp);
          }
        }
      } else {
        out.x = isbc ? (vi.x - p.bc_val[i]) : (au - p.alpha * p.f * p.fvec[i]);
        out.y = ap - p.bobs[i];
        part += out.x * out.x + out.y * out.y;
      }
      p.y[i] = out;
    }
  }
  if (p.partials) {
    const double r = lvpp_block_sum<256>(part, s_red);
    if (threadIdx.x == 0) p.partials[blockIdx.x] = r;
  }
}

// ------------------------------------------------------------------------------------------------
// The multigrid cycle's copy of the operator: one 16-byte record {column | bc bit, float(alpha K), float(M),
// float(D)} per sliced-ELL slot, so that a lane fetches a slot with ONE 128-bit load (a warp: 512 contiguous
// bytes per k) instead of four 32-bit streams, and 16 instead of 28 bytes per slot cross HBM.  The cycle is a
// preconditioner: its operator only has to be a fixed linear map, not the exact Jacobian; sums are still
// accumulated in fp64.  The record stream of the next U slots is requested before the gathers of the current U
// are issued (register double buffer): the two dependent memory phases (records, then v[col]) of consecutive
// groups overlap, which is what the four-stream version lacked (ncu: 4.9 TB/s, no unit above 41 %).
struct PackedOpArgs {
  int64_t Vown;
  const int64_t* slice_ptr;
  const uint4* P;        // [slots]
  const uint8_t* bc_flag;
  const double2* v;
  double2* y;
  int epi;               // EPI_NONE / EPI_RESID / EPI_JACOBI
  const double2* b;
  const double* binv;    // [Vown * 4]
  double omega;
  const int* skip;       // device flag, non-zero = the Krylov solve is over, do nothing (null: always run)
};

// loads as volatile asm: the compiler otherwise sinks the next group's record loads below the current group's
// arithmetic and serialises the gathers through one register quad (40 registers, one load in flight per thread)
__device__ __forceinline__ uint4 lvpp_ld_record(const uint4* ptr) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(ptr));
  return r;
}
__device__ __forceinline__ double2 lvpp_ld_pair(const double2* ptr) {
  double2 r;
  asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(ptr));
  return r;
}

template <int U, int MINB>
__global__ void __launch_bounds__(256, MINB) k_packed_op(PackedOpArgs p) {
  if (p.skip && *p.skip) return;
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x; i0 < p.Vown; i0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i0 + threadIdx.x;
    if (i >= p.Vown) continue;
    const int64_t s = i >> 5;
    const int64_t b0 = p.slice_ptr[s];
    const int w = (int)((p.slice_ptr[s + 1] - b0) >> 5);  // warp-uniform, >= 1 (every row holds its diagonal)
    const uint4* q = p.P + b0 + (i & 31);
    // slots past the end of the row re-read the last record (cache hit) and are zeroed: no predicated loads
    uint4 cur[U];
#pragma unroll
    for (int u = 0; u < U; ++u) cur[u] = lvpp_ld_record(q + (int64_t)min(u, w - 1) * LVPP_SLICE);
    double au = 0.0, ap = 0.0;
    for (int k0 = 0; k0 < w; k0 += U) {
      uint4 nxt[U];
#pragma unroll
      for (int u = 0; u < U; ++u) nxt[u] = lvpp_ld_record(q + (int64_t)min(k0 + U + u, w - 1) * LVPP_SLICE);
      double2 vj[U];
#pragma unroll
      for (int u = 0; u < U; ++u) vj[u] = lvpp_ld_pair(&p.v[cur[u].x & ~LVPP_COL_BC]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool live = k0 + u < w;
        const double kv = live ? (double)__uint_as_float(cur[u].y) : 0.0, mv = live ? (double)__uint_as_float(cur[u].z) : 0.0,
                     dv = live ? (double)__uint_as_float(cur[u].w) : 0.0;
        const double vx = (cur[u].x & LVPP_COL_BC) ? 0.0 : vj[u].x;
        au += kv * vx + mv * vj[u].y;
        ap += mv * vx - dv * vj[u].y;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) cur[u] = nxt[u];
    }
    const double2 vi = p.v[i];
    const bool isbc = p.bc_flag[i] != 0;
    double2 out;
    out.x = isbc ? vi.x : au;
    out.y = ap;
    if (p.epi != EPI_NONE) {
      const double2 bi = p.b[i];
      const double ru = bi.x - out.x, rp = bi.y - out.y;
      if (p.epi == EPI_RESID) {
        out.x = ru;
        out.y = rp;
      } else {
        const double* B = p.binv + 4 * i;
        out.x = vi.x + (isbc ? 1.0 : p.omega) * (B[0] * ru + B[1] * rp);
        out.y = vi.y + p.omega * (B[2] * ru + B[3] * rp);
      }
    }
    p.y[i] = out;
  }
}

// ------------------------------------------------------------------------------------------------
// The default records of the cycle (LVPP_MG_PACK=fp32 selects the ones above): bf16 values, one record per PAIR of consecutive slots of a row --
// a 128-bit load {col0, col1, bf16(alpha K0)|bf16(M0), bf16(alpha K1)|bf16(M1)} and a 32-bit load bf16(D0)|bf16(D1):
// 10 instead of 16 bytes per slot, and the node-block inverses in single precision (16 instead of 32 bytes per node):
// the fine-level sweep at n = 215 moves 2.15 instead of 3.22 GB.  bf16 keeps the exponent range
// of the single-precision records (D = int exp(psi) phi_i phi_j spans 26 orders of magnitude) and 8 bits of mantissa;
// on the CPU mirror of the cycle (tools/mg_precision.py) the Krylov iteration counts of the first four proximal steps
// are those of the fp64 cycle (360 against 359 in total on a 16^3 mesh) -- the cycle only has to be a fixed linear map
// close to J^-1.  Pair p of slice s sits at ((slice_ptr[s] + 32 s) >> 1) + 32 p + lane: rows of odd width get a zero
// second slot, no second pointer array is needed (16 records per slice of even width are left unused).
struct Packed2OpArgs {
  int64_t Vown;
  const int64_t* slice_ptr;
  const uint4* P2;
  const uint32_t* Pd;
  const uint8_t* bc_flag;
  const double2* v;
  double2* y;
  int epi;
  const double2* b;
  const float4* binv;    // [Vown] single-precision node-block inverses
  double omega;
  const int* skip;
};

__device__ __forceinline__ uint32_t lvpp_ld_u32(const uint32_t* ptr) {
  uint32_t r;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(ptr));
  return r;
}
__device__ __forceinline__ double lvpp_bf16_hi(uint32_t w) { return (double)__uint_as_float(w & 0xFFFF0000u); }
__device__ __forceinline__ double lvpp_bf16_lo(uint32_t w) { return (double)__uint_as_float(w << 16); }

template <int UP, int MINB>
__global__ void __launch_bounds__(256, MINB) k_packed2_op(Packed2OpArgs p) {
  if (p.skip && *p.skip) return;
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x; i0 < p.Vown; i0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i0 + threadIdx.x;
    if (i >= p.Vown) continue;
    const int64_t s = i >> 5;
    const int64_t b0 = p.slice_ptr[s];
    const int w = (int)((p.slice_ptr[s + 1] - b0) >> 5);
    const int np = (w + 1) >> 1;  // pair records of this row, >= 1
    const int64_t po = ((b0 + 32 * s) >> 1) + (i & 31);
    const uint4* q = p.P2 + po;
    const uint32_t* qd = p.Pd + po;
    uint4 cur[UP];
    uint32_t curd[UP];
#pragma unroll
    for (int u = 0; u < UP; ++u) {
      const int64_t off = (int64_t)min(u, np - 1) * LVPP_SLICE;
      cur[u] = lvpp_ld_record(q + off);
      curd[u] = lvpp_ld_u32(qd + off);
    }
    double au = 0.0, ap = 0.0;
    for (int k0 = 0; k0 < np; k0 += UP) {
      uint4 nxt[UP];
      uint32_t nxtd[UP];
#pragma unroll
      for (int u = 0; u < UP; ++u) {
        const int64_t off = (int64_t)min(k0 + UP + u, np - 1) * LVPP_SLICE;
        nxt[u] = lvpp_ld_record(q + off);
        nxtd[u] = lvpp_ld_u32(qd + off);
      }
      double2 va[UP], vb[UP];
#pragma unroll
      for (int u = 0; u < UP; ++u) {
        va[u] = lvpp_ld_pair(&p.v[cur[u].x & ~LVPP_COL_BC]);
        vb[u] = lvpp_ld_pair(&p.v[cur[u].y & ~LVPP_COL_BC]);
      }
#pragma unroll
      for (int u = 0; u < UP; ++u) {
        const bool live = k0 + u < np;  // records past the end of the row were re-reads of the last one
        const uint32_t km0 = live ? cur[u].z : 0u, km1 = live ? cur[u].w : 0u, dd = live ? curd[u] : 0u;
        const double vx0 = (cur[u].x & LVPP_COL_BC) ? 0.0 : va[u].x;
        const double vx1 = (cur[u].y & LVPP_COL_BC) ? 0.0 : vb[u].x;
        au += lvpp_bf16_hi(km0) * vx0 + lvpp_bf16_lo(km0) * va[u].y;
        ap += lvpp_bf16_lo(km0) * vx0 - lvpp_bf16_hi(dd) * va[u].y;
        au += lvpp_bf16_hi(km1) * vx1 + lvpp_bf16_lo(km1) * vb[u].y;
        ap += lvpp_bf16_lo(km1) * vx1 - lvpp_bf16_lo(dd) * vb[u].y;
      }
#pragma unroll
      for (int u = 0; u < UP; ++u) {
        cur[u] = nxt[u];
        curd[u] = nxtd[u];
      }
    }
    const double2 vi = p.v[i];
    const bool isbc = p.bc_flag[i] != 0;
    double2 out;
    out.x = isbc ? vi.x : au;
    out.y = ap;
    if (p.epi != EPI_NONE) {
      const double2 bi = p.b[i];
      const double ru = bi.x - out.x, rp = bi.y - out.y;
      if (p.epi == EPI_RESID) {
        out.x = ru;
        out.y = rp;
      } else {
        const float4 B = p.binv[i];
        out.x = vi.x + (isbc ? 1.0 : p.omega) * ((double)B.x * ru + (double)B.y * rp);
        out.y = vi.y + p.omega * ((double)B.z * ru + (double)B.w * rp);
      }
    }
    p.y[i] = out;
  }
}

static inline OpArgs lvpp_level_op(const lvpp_problem* h, const MgLevel& L) {
  OpArgs p;
  p.Vown = L.Vown; p.slice_ptr = L.slice_ptr; p.col = L.col;
  p.K = L.K; p.M = L.M; p.D = L.D; p.bc_flag = L.bc_flag; p.bc_val = nullptr;
  p.alpha = h->alpha; p.v = nullptr; p.xk = nullptr; p.bobs = nullptr; p.fvec = nullptr;
  p.f = 0.0; p.inv_scale = nullptr; p.skip_flag = nullptr; p.y = nullptr; p.partials = nullptr;
  p.epi = EPI_NONE; p.b = nullptr; p.binv = nullptr; p.omega = 1.0;
  return p;
}
