// admm_kernel.cuh -- shared device helpers + the one-instance-per-warp ADMM solver used by the TAIL kernel
// (per-instance KKT factor).  The main kernel (NI instances per warp) is admm_multi_kernel.cuh.
//
// Hand-written device code shared by every generated problem family; the generator emits
// only compile-time sizes (cpg_family.h), the blob-header struct (cpg_blob_layout.h) and the
// constants blob.  What it replaces, per instance, is the reference's generated
//   cpg_canonicalize_<id>()                       cvxpygen/utils.py:279-294      (a2)
//   osqp_update_data_vec / update bounds          osqp_sources/src/osqp.c:752-827 (a3)
//   osqp_solve main loop                          src/osqp.c:288-645              (a4)
//     compute_rhs / update_xz_tilde               src/auxil.c:161-183             (a5)
//     solve_linsys_qdldl -> QDLDL_solve           qdldl_interface.c:341-376, qdldl.c:236-281 (a6)
//     update_x / update_z+project / update_y      src/auxil.c:185-225, src/proj.c:4-14      (a7)
//     update_info, residuals, check_termination   src/auxil.c:240-359, 361-512, 564-629, 681-786 (a8)
//     compute_rho_estimate / adapt_rho            src/auxil.c:13-74               (a9, decision only)
//   store_solution / unscale_solution / obj       src/auxil.c:524-562, src/scaling.c:177-192 (a11)
//   cpg_retrieve_prim / dual / info               cvxpygen/utils.py:950-985       (a12)
//
// Data layout
//   * constants blob (factor schedule, scaled A/P for the residuals, scalings, maps) staged ONCE per
//     CTA global->shared with TMA bulk copies (cp.async.bulk + mbarrier complete_tx);
//   * per-instance ADMM state x, z, y, q, l, u lives in REGISTERS, element i of a vector on lane i%32;
//   * the only per-warp shared memory is the (n+m)-double work vector w of the KKT solve;
//   * HBM is touched only to read the instance's parameters and to write its solution rows
//     (coalesced: a warp writes consecutive doubles of one row).
//   * persistent CTAs: warps pull instance indices from a global counter, so instances that stop
//     at iteration 25 do not leave lanes idle while a neighbour runs to 100.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "cpg_blob_layout.h"

namespace cpgb200 {

constexpr int LANES = 32;
constexpr unsigned FULL = 0xffffffffu;
constexpr double OSQP_INFTY = 1e30;
constexpr double MIN_SCALING = 1e-4;
constexpr double RHO_MIN = 1e-6, RHO_MAX = 1e6, RHO_TOL = 1e-4, RHO_EQ_FACTOR = 1e3;
constexpr double DIVISION_TOL = 1.0 / OSQP_INFTY;

// status codes = OSQP's (osqp_sources/include/constants.h:18-30) plus one internal hand-off code
enum : int { ST_SOLVED = 1, ST_SOLVED_INACC = 2, ST_PINF_INACC = 3, ST_DINF_INACC = 4, ST_MAXITER = -2,
             ST_PINF = -3, ST_DINF = -4, ST_NONCVX = -7, ST_UNSOLVED = -10, ST_HANDOFF = -100 };

struct Settings {          // cvxpygen's OSQP settings table, cvxpygen/solvers/osqp.py:102-115
  int max_iter, check_termination, scaled_termination, warm_start;
  int adaptive_rho, adaptive_rho_interval, scaling, pad;
  double eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, alpha, adaptive_rho_tolerance;
};

struct BatchIO {
  const double* params;    // (B, npb) row-major: the batched user parameters of each instance
  const double* x0;        // optional warm start, canonical, UNscaled: (B, n) / (B, m)
  const double* y0;
  double* prim;            // (B, n_prim) user-level primal variables, concatenated in declaration order
  double* dual;            // (B, n_dual) user-level constraint duals
  double* sol_x;           // optional canonical solution (B, n) / (B, m)   (gradient path needs it)
  double* sol_y;
  double* obj_val; int* iter; int* status; double* pri_res; double* dua_res;   // CPG_Info, SoA
  unsigned int* work_counter;   // persistent-CTA instance queue
  int* tail_count;              // instances handed to the refactorisation kernel
  int* tail_ids;
  double* tail_state;           // per handed-off instance: x(n) z(m) y(m) rho_new iter   (scaled iterates)
  int B;
  int tail_capacity;
  double* ws;                   // warm start only: per instance z0 = A x0 (m) | y0 / rho (m), the start point of the main kernel
};

// ---------------------------------------------------------------- TMA bulk copy + mbarrier helpers
#ifdef CPG_SIMT_HOST_EMU
// host build of the test suite (tests/emu/simt): the staging copy is a memcpy by the issuing thread, the wait a block barrier
__device__ __forceinline__ void mbar_init(uint64_t*, int) {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t*, uint32_t) {}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t*) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void mbar_wait(uint64_t*, uint32_t) { __syncthreads(); }
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#endif

// ---------------------------------------------------------------- warp reductions
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// ---------------------------------------------------------------- ELL sparse dot products
// row-blocked ELL: table entry = {K, f64 offset, u16 offset}; entry k of lane's row at [k*32 + lane]
__device__ __forceinline__ double ell_dot(const int* tab, const double* F64, const uint16_t* U16,
                                          const double* vec, int lane) {
  const int K = tab[0];
  const double* v = F64 + tab[1] + lane;
  const uint16_t* c = U16 + tab[2] + lane;
  double a0 = 0.0, a1 = 0.0;
  int k = 0;
  for (; k + 1 < K; k += 2) {
    a0 = fma(v[k * LANES], vec[c[k * LANES]], a0);
    a1 = fma(v[(k + 1) * LANES], vec[c[(k + 1) * LANES]], a1);
  }
  if (k < K) a0 = fma(v[k * LANES], vec[c[k * LANES]], a0);
  return a0 + a1;
}

// ---------------------------------------------------------------- tail path: per-instance numeric LDL' (a9)
// Tables built by offline/refactor.py; they live in GLOBAL memory (L2-resident, shared by all tail warps).
struct TailView {
  const CpgTailHeader* H;
  const int* I32;
  const double* F64;
  const uint16_t* U16;
};
__device__ __forceinline__ TailView make_tail_view(const uint8_t* blob) {
  TailView tv;
  tv.H = reinterpret_cast<const CpgTailHeader*>(blob);
  tv.I32 = reinterpret_cast<const int*>(blob + tv.H->off_i32);
  tv.F64 = reinterpret_cast<const double*>(blob + tv.H->off_f64);
  tv.U16 = reinterpret_cast<const uint16_t*>(blob + tv.H->off_u16);
  return tv;
}

// Per-level pointers of the factorisation (level_ptr | coloured round_ptr | scale_ptr, n_levels + 1 entries each): a compile-time
// table in constant memory when the family header carries it -- a uniform LDC instead of a dependent global load in front of
// every level's loops (pointer chasing behind an L2 latency, 134 levels x 3 loops for mpc_ltv_12_4_10).
#ifdef CPG_FAM_TAIL_LEVELS
__constant__ const int kTailLevels[] = CPG_FAM_TAIL_LEVELS;
#define CPG_TAIL_LEVEL_PTR(tv, which, off) (kTailLevels + (which) * ((tv).H->n_levels + 1))
#else
#define CPG_TAIL_LEVEL_PTR(tv, which, off) ((tv).I32 + (off))
#endif
// Forms of the factorisation's update phase (all checked on the SIMT emulator, tests/test_simt_emulation.py):
//   CPG_TAIL_FACTOR_FORM 2 (default)  COLOURED ROUNDS: the ops of a level are dealt offline to rounds of 32 with pairwise distinct
//       targets (offline/refactor.py: greedy, reaches ceil(ops / 32) rounds on the MPC families); a round is a plain read-modify-
//       write per lane + __syncwarp -- no atomics, one fixed order per target (bit-reproducible), balanced lanes;
//   0  push form with shared-memory atomicAdd(double) -- a CAS loop on sm_100 (13 % of the matrix-parameter kernel's samples,
//       profiles/r2_matpar_v7_ncu_summary.md), run-to-run differences in the last bits;
//   1  owner-writes: a target's ops summed by ONE lane (deterministic, but unbalanced: measured 1 - 8 % slower than the atomics,
//       profiles/r2_tail_gather_factor_ab.jsonl).  (CPG_TAIL_GATHER_FACTOR=1 is the older spelling of form 1.)
#ifndef CPG_TAIL_FACTOR_FORM
#if defined(CPG_TAIL_GATHER_FACTOR) && CPG_TAIL_GATHER_FACTOR
#define CPG_TAIL_FACTOR_FORM 1
#else
#define CPG_TAIL_FACTOR_FORM 2
#endif
#endif
// Numeric factorisation of K(rho_vec) on the family's symbolic pattern (role of QDLDL_factor, qdldl.c:72-233,
// after update_KKT_param2, kkt.c:214-222).  S holds K's lower triangle in slot order with -1/rho_vec already
// written; on exit S[j] = 1/D_j and the other slots hold L.  Right-looking, one elimination-tree level at a time.
__device__ __forceinline__ void tail_factor(const TailView& tv, double* S, int lane) {
  const int nl = tv.H->n_levels;
  const int* lp = CPG_TAIL_LEVEL_PTR(tv, 0, tv.H->i_level_ptr);
  const int* sp = CPG_TAIL_LEVEL_PTR(tv, 2, tv.H->i_scale_ptr);
  const uint16_t* lc = tv.U16 + tv.H->h_level_cols;
  const ushort2* scl = reinterpret_cast<const ushort2*>(tv.U16 + tv.H->h_scale);
#if CPG_TAIL_FACTOR_FORM == 2
  const int* rp = CPG_TAIL_LEVEL_PTR(tv, 1, tv.H->i_cround_ptr);
  const ushort4* cops = reinterpret_cast<const ushort4*>(tv.U16 + tv.H->h_cops) + lane;
#elif CPG_TAIL_FACTOR_FORM == 1
  const int* gtp = tv.I32 + tv.H->i_gtgt_ptr;
  const int* gsg = tv.I32 + tv.H->i_gseg;
  const ushort4* gops = reinterpret_cast<const ushort4*>(tv.U16 + tv.H->h_gops);
#else
  const int* op = tv.I32 + tv.H->i_op_ptr;
  const ushort4* ops = reinterpret_cast<const ushort4*>(tv.U16 + tv.H->h_ops);
#endif
  for (int lv = 0; lv < nl; ++lv) {
    for (int c = lp[lv] + lane; c < lp[lv + 1]; c += LANES) { const int j = lc[c]; S[j] = 1.0 / S[j]; }
    __syncwarp();
#if CPG_TAIL_FACTOR_FORM == 2
    // rounds of one SYNC GROUP touch pairwise distinct targets: no barrier between them (bit 15 of the op's last field marks the
    // round a barrier follows; a level that eliminates one column is a single group); the table words run two rounds ahead
    {
      const int r1 = rp[lv + 1];
      int r = rp[lv];
      if (r < r1) {
        ushort4 q = __ldg(cops + (size_t)r * LANES);
        ushort4 qn = (r + 1 < r1) ? __ldg(cops + (size_t)(r + 1) * LANES) : q;
        for (; r < r1; ++r) {
          const ushort4 qnn = (r + 2 < r1) ? __ldg(cops + (size_t)(r + 2) * LANES) : qn;
          S[q.x] -= S[q.y] * S[q.z] * S[q.w & 0x7fffu];      // (a padding op subtracts 0 from the lane's own dummy slot beyond n_slots)
          if (q.w & 0x8000u) __syncwarp();
          q = qn; qn = qnn;
        }
      }
    }
#elif CPG_TAIL_FACTOR_FORM == 1
    // owner-writes form: a target's ops are contiguous and summed by ONE lane in table order (no atomics, deterministic)
    for (int ti = gtp[lv] + lane; ti < gtp[lv + 1]; ti += LANES) {
      const int o0 = __ldg(gsg + ti), o1 = __ldg(gsg + ti + 1);
      ushort4 q = __ldg(gops + o0);
      double acc = S[q.y] * S[q.z] * S[q.w];
      for (int o = o0 + 1; o < o1; ++o) { q = __ldg(gops + o); acc += S[q.y] * S[q.z] * S[q.w]; }
      S[q.x] -= acc;
    }
    __syncwarp();
#else
    for (int o = op[lv] + lane; o < op[lv + 1]; o += LANES) {
      const ushort4 q = __ldg(ops + o);
      atomicAdd(&S[q.x], -(S[q.y] * S[q.z] * S[q.w]));
    }
    __syncwarp();
#endif
    for (int o = sp[lv] + lane; o < sp[lv + 1]; o += LANES) { const ushort2 q = __ldg(scl + o); S[q.x] *= S[q.y]; }
    __syncwarp();
  }
}

// entries of a tile: one 32-bit word per (k, lane) = slot | position << 16, lane-interleaved (offline/blob.py:pack_tail_blob);
// families with fewer than 8192 slots carry both fields as BYTE offsets (CPG_FAM_TAIL_WORD_SHIFT = 3): no shift per operand
#ifndef CPG_FAM_TAIL_WORD_SHIFT
#define CPG_FAM_TAIL_WORD_SHIFT 0
#endif
__device__ __forceinline__ double tail_entry_fma(const double* S, const double* w, unsigned e, double a) {
#if CPG_FAM_TAIL_WORD_SHIFT == 3
  return fma(*reinterpret_cast<const double*>(reinterpret_cast<const char*>(S) + (e & 0xffffu)),
             *reinterpret_cast<const double*>(reinterpret_cast<const char*>(w) + (e >> 16)), a);
#else
  return fma(S[e & 0xffffu], w[e >> 16], a);
#endif
}
// The first TAIL_PRE words per lane of a tile are fetched ONE TILE AHEAD (tail_prefetch, issued before the previous tile's
// in-register sweep): the words are static addresses behind an L2 latency -- shared memory is carved out for the factors, L1
// keeps ~28 KB that the words of one solve (29 KB for mpc_ltv_12_4_10) flush -- and were the kernel's largest stall
// (long_scoreboard 2.9 cycles per issue, profiles/r2_matpar_v5_ncu_summary.md).  Slots beyond the tile's K hold `zero_word`
// (S[zero slot] * w[0] = 0).  Entry k accumulates into chain k mod 4.
// Depth measured on B200 (profiles/r2_tail_prefetch_ab.jsonl): the matrix-parameter kernel gains 7.7 % from 12 words instead of 8
// (86.3 -> 79.7 ms per 20 000; 4: 88.7, 16: 81.7, 20: 79.6), the tail and backward kernels of shared-matrix families do not (+-1 %).
#ifndef CPG_TAIL_PRE
#if defined(CPG_FAM_MATPAR) && CPG_FAM_MATPAR
#define CPG_TAIL_PRE 12
#else
#define CPG_TAIL_PRE 8
#endif
#endif
constexpr int TAIL_PRE = CPG_TAIL_PRE;
__device__ __forceinline__ void tail_prefetch(unsigned (&pre)[TAIL_PRE], const int* h, const int* __restrict__ I32, int lane,
                                              unsigned zero_word) {
  const unsigned* wd = reinterpret_cast<const unsigned*>(I32) + h[0] + lane;
  const int K = h[2];
#pragma unroll
  for (int u = 0; u < TAIL_PRE; ++u) pre[u] = (u < K) ? __ldg(wd + u * LANES) : zero_word;
}
__device__ __forceinline__ double slot_tile_acc(const int* h, const int* __restrict__ I32, const double* S,
                                                const double* w, int lane, const unsigned (&pre)[TAIL_PRE]) {
  const unsigned* wd = reinterpret_cast<const unsigned*>(I32) + h[0] + lane;
  const int K = h[2];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
  for (int u = 0; u < TAIL_PRE; u += 4) {
    a0 = tail_entry_fma(S, w, pre[u], a0);
    a1 = tail_entry_fma(S, w, pre[u + 1], a1);
    a2 = tail_entry_fma(S, w, pre[u + 2], a2);
    a3 = tail_entry_fma(S, w, pre[u + 3], a3);
  }
  int k = TAIL_PRE;
  for (; k + 3 < K; k += 4) {
    const unsigned e0 = __ldg(wd + k * LANES), e1 = __ldg(wd + (k + 1) * LANES);
    const unsigned e2 = __ldg(wd + (k + 2) * LANES), e3 = __ldg(wd + (k + 3) * LANES);
    a0 = tail_entry_fma(S, w, e0, a0);
    a1 = tail_entry_fma(S, w, e1, a1);
    a2 = tail_entry_fma(S, w, e2, a2);
    a3 = tail_entry_fma(S, w, e3, a3);
  }
  if (k < K) {
    a0 = tail_entry_fma(S, w, __ldg(wd + k * LANES), a0);
    if (k + 1 < K) a1 = tail_entry_fma(S, w, __ldg(wd + (k + 1) * LANES), a1);
    if (k + 2 < K) a2 = tail_entry_fma(S, w, __ldg(wd + (k + 2) * LANES), a2);
  }
  double acc = (a0 + a1) + (a2 + a3);
  for (int o = 16; o >= h[3]; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
  return acc;
}

// Resolve the dependencies INSIDE a group tile in registers: lane t holds the value of row t of the tile; rows are
// finalised one after the other (ascending for L, descending for L'), each broadcast with one shuffle and applied with
// one FMA by the lanes that depend on it.  Coefficients are fetched eight rows ahead so that their shared-memory
// latency overlaps the sweep.  inside[j*32 + t] = slot of the coefficient coupling row t to row j (zero slot if none).
template <bool ASCENDING>
__device__ __forceinline__ double group_sweep(double val, const uint16_t* __restrict__ inside, const double* __restrict__ S,
                                              const int nrows, const int lane) {
  const int nblk = (nrows + 7) >> 3;
  for (int blk = 0; blk < nblk; ++blk) {
    double cf[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int jj = blk * 8 + u;
      const int j = ASCENDING ? jj : nrows - 1 - jj;
      cf[u] = (jj < nrows) ? S[inside[j * LANES + lane]] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int jj = blk * 8 + u;
      if (jj < nrows) {
        const int j = ASCENDING ? jj : nrows - 1 - jj;
        const double vj = __shfl_sync(FULL, val, j);
        const bool dep = ASCENDING ? (lane > j) : (lane < j);
        if (dep) val = fma(-cf[u], vj, val);
      }
    }
  }
  return val;
}

// The same sweep when the group's couplings sit in a packed strict-lower triangle Sg[tri(j) + t] (j = row being swept,
// t > j its dependents; offline/refactor.py:tri_offset): no index table -- an ascending sweep reads consecutive addresses
// across the lanes, a descending one reads column t of the triangle at a per-lane base.
// tri(j) = j (g - 1) - j (j - 1) / 2 - (j + 1);   coupling (t, j), t > j, at tri(j) + t;   tri(j + 1) - tri(j) = g - 2 - j
//
// Profile of the first version (profiles/r1_matpar_v2, source page): these sweeps were HALF of the matrix-parameter kernel's
// instructions -- 18 per row (per-row predicates, address arithmetic, divergence bookkeeping) around the 4 that do the work
// (two SHFL halves, one LDS, one DFMA).  Now: the coefficient loads are unconditional (a lane that does not depend on row j reads
// a valid neighbouring slot and discards it -- the update is a predicated DFMA), and a full group (g = 32, the chain groups of the
// MPC families) runs fully unrolled with every row index, shuffle source and triangle offset an immediate.
// val -= cf * vj on the lanes where `on` holds: one ISETP (folded by the compiler when `on` compares against an immediate) and a
// PREDICATED DFMA.  Written as `if (on) val = fma(..)` the compiler computes the product everywhere and selects with two FSEL,
// and packs the 32 predicates of an unrolled sweep into a register with a LOP3 each (profiles/r2_matpar_v5: 10 instructions
// per row instead of 5).
template <bool GT>      // GT: lanes above row j take the update (ascending sweep); else the lanes below it
__device__ __forceinline__ void fnma_if(double& val, const double cf, const double vj, const int lane, const int j) {
#ifdef CPG_SIMT_HOST_EMU
  if (GT ? lane > j : lane < j) val = fma(-cf, vj, val);
#else
  if (GT) asm("{\n .reg .pred p;\n .reg .f64 n;\n setp.gt.s32 p, %3, %4;\n neg.f64 n, %1;\n @p fma.rn.f64 %0, n, %2, %0;\n}"
              : "+d"(val) : "d"(cf), "d"(vj), "r"(lane), "r"(j));
  else asm("{\n .reg .pred p;\n .reg .f64 n;\n setp.lt.s32 p, %3, %4;\n neg.f64 n, %1;\n @p fma.rn.f64 %0, n, %2, %0;\n}"
           : "+d"(val) : "d"(cf), "d"(vj), "r"(lane), "r"(j));
#endif
}

constexpr __host__ __device__ int tri32(int j) { return j * 31 - (j * (j - 1)) / 2 - (j + 1); }

template <bool ASCENDING>
__device__ __forceinline__ double group_sweep_dense32(double val, const double* __restrict__ Sg, const int lane) {
  if (ASCENDING) {
    const double* p = Sg + lane;                       // coupling (lane, j) at Sg[tri(j) + lane]; Sg[-1] exists (slot >= nk)
#pragma unroll
    for (int blk = 0; blk < 4; ++blk) {
      double cf[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) cf[u] = p[tri32(blk * 8 + u)];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = blk * 8 + u;
        const double vj = __shfl_sync(FULL, val, j);
        fnma_if<true>(val, cf[u], vj, lane, j);
      }
    }
  } else {
    const double* p = Sg + (lane * 31 - (lane * (lane - 1)) / 2 - (lane + 1));   // coupling (j, lane), j > lane, at Sg[tri(lane) + j]
#pragma unroll
    for (int blk = 0; blk < 4; ++blk) {
      double cf[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) cf[u] = p[31 - blk * 8 - u];      // lane 31 reads up to 31 slots past its (empty) column: inside S
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = 31 - blk * 8 - u;
        const double vj = __shfl_sync(FULL, val, j);
        fnma_if<false>(val, cf[u], vj, lane, j);
      }
    }
  }
  return val;
}

template <bool ASCENDING>
__device__ __forceinline__ double group_sweep_dense(double val, const double* __restrict__ Sg, const int g, const int lane) {
  if (g == 32) return group_sweep_dense32<ASCENDING>(val, Sg, lane);
  const int tl = lane < g ? lane : g - 1;              // lanes beyond the group compute on a copy of the last row (never read)
  if (ASCENDING) {
    const double* p = Sg - 1 + tl;                     // tri(0) + lane
    int step = g - 2;                                  // tri(j + 1) - tri(j) at j = 0
    int j = 0;
    for (; j + 4 <= g; j += 4) {
      double cf[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { cf[u] = *p; p += step; --step; }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double vj = __shfl_sync(FULL, val, j + u);
        fnma_if<true>(val, cf[u], vj, lane, j + u);
      }
    }
    for (; j < g; ++j) {
      const double cf = *p; p += step; --step;
      const double vj = __shfl_sync(FULL, val, j);
      fnma_if<true>(val, cf, vj, lane, j);
    }
  } else {
    const double* p = Sg + (tl * (g - 1) - (tl * (tl - 1)) / 2 - (tl + 1)) + (g - 1);     // tri(lane) + j at j = g - 1
    int j = g - 1;
    for (; j >= 3; j -= 4) {
      double cf[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) cf[u] = p[-u];
      p -= 4;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double vj = __shfl_sync(FULL, val, j - u);
        fnma_if<false>(val, cf[u], vj, lane, j - u);
      }
    }
    for (; j >= 0; --j) {
      const double cf = *p; --p;
      const double vj = __shfl_sync(FULL, val, j);
      fnma_if<false>(val, cf, vj, lane, j);
    }
  }
  return val;
}

// Tile headers of the triangular solves (8 ints per tile, offline/blob.py:pack_tail_blob).  They are the same for every warp and
// every instance, so a family's generated header carries them as a compile-time table that lands in CONSTANT memory: one uniform
// LDC per field instead of dependent global loads behind an L1 that the entry words keep flushing (the header / row-list loads
// were the three largest stall sites of the first version).  Code generated before this table existed reads them from the blob.
#ifdef CPG_FAM_TAIL_TILES
__constant__ const int kTailTiles[] = CPG_FAM_TAIL_TILES;
#define CPG_TAIL_TILE_PTR(tv) (kTailTiles)
#else
#define CPG_TAIL_TILE_PTR(tv) ((tv).I32 + (tv).H->i_tiles)
#endif

// K x = b with the per-instance factor: grouped level-scheduled L solve, D^{-1}, L' solve (QDLDL_solve, qdldl.c:269-281)
__device__ __forceinline__ void tail_solve(const TailView& tv, const double* S, double* w, int lane) {
  const int* T = CPG_TAIL_TILE_PTR(tv);
  const int nf = tv.H->n_fwd_tiles, nt = nf + tv.H->n_bwd_tiles, nk = tv.H->nk;
  const unsigned zero_word = (unsigned)(tv.H->n_slots - 1) << CPG_FAM_TAIL_WORD_SHIFT;
  unsigned pre[TAIL_PRE];
  if (nt > 0) tail_prefetch(pre, T, tv.I32, lane, zero_word);
  for (int t = 0; t < nt; ++t) {
    if (t == nf) {
      for (int i = lane; i < nk; i += LANES) w[i] *= S[i];
      __syncwarp();
    }
    const int* h = T + 8 * t;
    const int row = __ldg(tv.U16 + h[5] + lane);        // issued before the entry words: its latency hides behind them
    const double acc = slot_tile_acc(h, tv.I32, S, w, lane, pre);
    if (t + 1 < nt) tail_prefetch(pre, h + 8, tv.I32, lane, zero_word);     // in flight during this tile's sweep
    const int nrows = h[4];
    double val = w[row] - acc;
    if (h[7] == 2) val = (t < nf) ? group_sweep_dense<true>(val, S + h[6], nrows, lane)
                                  : group_sweep_dense<false>(val, S + h[6], nrows, lane);
    else if (h[7]) val = (t < nf) ? group_sweep<true>(val, tv.U16 + h[6], S, nrows, lane)
                                  : group_sweep<false>(val, tv.U16 + h[6], S, nrows, lane);
    __syncwarp();
    if (lane < nrows) w[row] = val;
    __syncwarp();
  }
  if (nt == nf) { for (int i = lane; i < nk; i += LANES) w[i] *= S[i]; __syncwarp(); }
}

// ---------------------------------------------------------------- matrix-parameter path
// Tables of offline/blob.py:pack_matpar_blob in GLOBAL memory + the per-warp shared arrays that hold ONE instance's
// scaled matrices and scalings (filled by matpar_prepare, matpar_kernel.cuh).
struct MatView {
  const CpgMatHeader* H;
  const int* I32;
  const double* F64;
  const uint16_t* U16;
};
__device__ __forceinline__ MatView make_mat_view(const uint8_t* blob) {
  MatView mv;
  mv.H = reinterpret_cast<const CpgMatHeader*>(blob);
  mv.I32 = reinterpret_cast<const int*>(blob + mv.H->off_i32);
  mv.F64 = reinterpret_cast<const double*>(blob + mv.H->off_f64);
  mv.U16 = reinterpret_cast<const uint16_t*>(blob + mv.H->off_u16);
  return mv;
}
struct MatCtx {
  MatView mv;
  double* Av;             // nnzA + 1 scaled entries of A in CSC order (last = 0: padding target of the index tables)
  double* Pv;             // nnzP + 1 scaled entries of the upper triangle of P
  double* D; double* Dinv; double* E; double* Einv;   // this instance's equilibration (scaling.c:44-156)
  double c, cinv;
};
// row-blocked ELL of INDICES: entry k of lane's row is value vals[idx[k*32+lane]] times vec[col[k*32+lane]]
__device__ __forceinline__ double ellx_dot(const int* __restrict__ tab, const uint16_t* __restrict__ U16,
                                           const double* vals, const double* vec, int lane) {
  const int K = __ldg(tab);
  const uint16_t* ix = U16 + __ldg(tab + 1) + lane;
  const uint16_t* c = U16 + __ldg(tab + 2) + lane;
  double a0 = 0.0, a1 = 0.0;
  int k = 0;
  for (; k + 1 < K; k += 2) {
    a0 = fma(vals[__ldg(ix + k * LANES)], vec[__ldg(c + k * LANES)], a0);
    a1 = fma(vals[__ldg(ix + (k + 1) * LANES)], vec[__ldg(c + (k + 1) * LANES)], a1);
  }
  if (k < K) a0 = fma(vals[__ldg(ix + k * LANES)], vec[__ldg(c + k * LANES)], a0);
  return a0 + a1;
}
// CPG_EQ_UNROLL = 1 keeps four (index, value) loads of the equilibration's norm / scaling loops in flight per lane.  Measured on
// B200 it is SLOWER (mpc_ltv_12_4_10, 20 000 instances: 105.6 vs 94.8 ms, profiles/r2_matpar_factor_unroll_ab.jsonl) -- the kernel
// sits at 255 registers and pays for the extra live values elsewhere -- so it is off.
#ifndef CPG_EQ_UNROLL
#define CPG_EQ_UNROLL 0
#endif
__device__ __forceinline__ double ellx_absmax(const int* __restrict__ tab, const uint16_t* __restrict__ U16,
                                              const double* vals, int lane) {
  const int K = __ldg(tab);
  const uint16_t* ix = U16 + __ldg(tab + 1) + lane;
  // index and value both sit behind an L2 latency (the entries of A live in the warp's global scratch slice): four independent
  // chains in flight; a maximum does not depend on the order
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int k = 0;
#if CPG_EQ_UNROLL
  for (; k + 3 < K; k += 4) {
    const int i0 = __ldg(ix + k * LANES), i1 = __ldg(ix + (k + 1) * LANES), i2 = __ldg(ix + (k + 2) * LANES), i3 = __ldg(ix + (k + 3) * LANES);
    const double v0 = vals[i0], v1 = vals[i1], v2 = vals[i2], v3 = vals[i3];
    a0 = fmax(a0, fabs(v0)); a1 = fmax(a1, fabs(v1)); a2 = fmax(a2, fabs(v2)); a3 = fmax(a3, fabs(v3));
  }
#endif
  for (; k < K; ++k) a0 = fmax(a0, fabs(vals[__ldg(ix + k * LANES)]));
  return fmax(fmax(a0, a1), fmax(a2, a3));
}

struct TailArgs {
  TailView tv;
  double* S;              // per-warp shared memory, n_slots doubles
  const double* state;    // x(n) z(m) y(m) rho iter of the handed-off instance
  MatCtx* mc;             // MODE 2 only
};

// ---------------------------------------------------------------- per-instance solver
template <class Fam>
struct Instance {
  static constexpr int N = Fam::N, M = Fam::M, NXL = (N + 31) / 32, NZL = (M + 31) / 32;
  static constexpr int NZLs = NZL > 0 ? NZL : 1;

  double x[NXL], q[NXL];
  double z[NZLs], y[NZLs], l[NZLs], u[NZLs];
  double dx[NXL], dy[NZLs];
  unsigned eqmask, loosemask;     // bit k: constraint row lane+32k is an equality / a loose row

  // residual bookkeeping of the last update_info
  double pri_res, dua_res, xPx, qx;
  double nrm_z, nrm_Ax, nrm_q, nrm_Aty, nrm_Px;                 // termination norms (unscaled unless scaled_termination)
  double s_rp, s_rd, s_z, s_Ax, s_q, s_Aty, s_Px;               // SCALED norms for compute_rho_estimate

  __device__ __forceinline__ double rho_of(int k, double rho_in, double rho_eq) const {
    return ((loosemask >> k) & 1u) ? RHO_MIN : (((eqmask >> k) & 1u) ? rho_eq : rho_in);
  }
};

// MODE 0: family factor from the blob (legacy one-instance-per-warp path); MODE 1: tail kernel, per-instance factor of
// K(rho_vec) with the family's matrices; MODE 2: matrix-parameter kernel, per-instance matrices, scalings and factor.
template <class Fam, int MODE>
__device__ void solve_instance(const CpgBlobHeader* __restrict__ H, const int* __restrict__ I32,
                               const double* __restrict__ F64, const uint16_t* __restrict__ U16,
                               double* __restrict__ w, const int lane, const int b,
                               const BatchIO& io, const Settings& st, const TailArgs* ta) {
  using I = Instance<Fam>;
  constexpr int N = I::N, M = I::M, NXL = I::NXL, NZL = I::NZL;
  constexpr bool TAIL = MODE >= 1, MATPAR = MODE == 2;
  I s;
  const double* Dv = MATPAR ? ta->mc->D : F64 + H->f_D;  const double* Dinv = MATPAR ? ta->mc->Dinv : F64 + H->f_Dinv;
  const double* Ev = MATPAR ? ta->mc->E : F64 + H->f_E;  const double* Einv = MATPAR ? ta->mc->Einv : F64 + H->f_Einv;
  const double c = MATPAR ? ta->mc->c : H->c, cinv = MATPAR ? ta->mc->cinv : H->cinv, sigma = H->sigma, alpha = st.alpha;
  // products with A, A' and the full symmetric P: family constants (blob ELL) or this instance's values (index ELL)
  auto dotA = [&](int k, const double* vec) -> double {
    if constexpr (MATPAR) return ellx_dot(ta->mc->mv.I32 + ta->mc->mv.H->i_ixA + 3 * k, ta->mc->mv.U16, ta->mc->Av, vec, lane);
    else return ell_dot(I32 + H->i_ellA + 3 * k, F64, U16, vec, lane);
  };
  auto dotAt = [&](int k, const double* vec) -> double {
    if constexpr (MATPAR) return ellx_dot(ta->mc->mv.I32 + ta->mc->mv.H->i_ixAt + 3 * k, ta->mc->mv.U16, ta->mc->Av, vec, lane);
    else return ell_dot(I32 + H->i_ellAt + 3 * k, F64, U16, vec, lane);
  };
  auto dotP = [&](int k, const double* vec) -> double {
    if constexpr (MATPAR) return ellx_dot(ta->mc->mv.I32 + ta->mc->mv.H->i_ixP + 3 * k, ta->mc->mv.U16, ta->mc->Pv, vec, lane);
    else return ell_dot(I32 + H->i_ellP + 3 * k, F64, U16, vec, lane);
  };
  const bool unscale = st.scaling && !st.scaled_termination;

  // ---- a1/a2/a3: canonicalise the instance's vectors and scale them (q <- c D q ; l,u <- E l, E u)
  const double* th = io.params + (size_t)b * H->npb;
  int px[NXL], pz[I::NZLs];
  bool type_mismatch = false;
  s.eqmask = 0u; s.loosemask = 0u;
#pragma unroll
  for (int k = 0; k < NXL; ++k) {
    const int i = lane + 32 * k;
    s.q[k] = 0.0; s.x[k] = 0.0; s.dx[k] = 0.0; px[k] = 0;
    if (i < N) {
      const int* tab = I32 + H->i_ellMq + 3 * k;
      double acc = F64[H->f_qbase + i];
      const int K = tab[0];
      for (int kk = 0; kk < K; ++kk)
        acc = fma(F64[tab[1] + kk * LANES + lane], __ldg(th + U16[tab[2] + kk * LANES + lane]), acc);
      s.q[k] = (Dv[i] * acc) * c;
      px[k] = U16[H->h_pinvx + i];
    }
  }
#pragma unroll
  for (int k = 0; k < NZL; ++k) {
    const int j = lane + 32 * k;
    s.l[k] = 0.0; s.u[k] = 0.0; s.z[k] = 0.0; s.y[k] = 0.0; s.dy[k] = 0.0; pz[k] = 0;
    if (j < M) {
      const int* tl = I32 + H->i_ellMl + 3 * k;
      const int* tu = I32 + H->i_ellMu + 3 * k;
      double al = F64[H->f_lbase + j], au = F64[H->f_ubase + j];
      for (int kk = 0; kk < tl[0]; ++kk)
        al = fma(F64[tl[1] + kk * LANES + lane], __ldg(th + U16[tl[2] + kk * LANES + lane]), al);
      for (int kk = 0; kk < tu[0]; ++kk)
        au = fma(F64[tu[1] + kk * LANES + lane], __ldg(th + U16[tu[2] + kk * LANES + lane]), au);
      al = fmin(fmax(al, -OSQP_INFTY), OSQP_INFTY);
      au = fmin(fmax(au, -OSQP_INFTY), OSQP_INFTY);
      s.l[k] = Ev[j] * al; s.u[k] = Ev[j] * au;
      pz[k] = U16[H->h_pinvz + j];
      // constraint type on the scaled bounds (set_rho_vec / update_rho_vec, auxil.c:76-142)
      const bool loose = (s.l[k] < -OSQP_INFTY * MIN_SCALING) && (s.u[k] > OSQP_INFTY * MIN_SCALING);
      const bool eq = !loose && (s.u[k] - s.l[k] < RHO_TOL);
      const int ct = loose ? 0 : (eq ? 2 : 1);
      type_mismatch |= (ct != (int)U16[H->h_ctype + j]);
      if (eq) s.eqmask |= 1u << k;
      if (loose) s.loosemask |= 1u << k;
    }
  }
  type_mismatch = __any_sync(FULL, type_mismatch);

  double rho_in = H->rho, rho_eq = RHO_EQ_FACTOR * H->rho;
  double rinv_in = 1.0 / rho_in, rinv_eq = 1.0 / rho_eq;
  const double rinv_loose = 1.0 / RHO_MIN;
  auto rinv_of = [&](int k) -> double {
    return ((s.loosemask >> k) & 1u) ? rinv_loose : (((s.eqmask >> k) & 1u) ? rinv_eq : rinv_in);
  };

  // ---- cold start (auxil.c:155-159) or warm start (osqp.c:929-953: x <- Dinv x, y <- c Einv y, z <- A x)
  int it0 = 0;
  // (TAIL) build this instance's own factor: S <- K(rho_vec) on the symbolic pattern, then numeric LDL'
  auto refactor = [&]() {
    if (TAIL) {
      const TailView& tv = ta->tv;
      double* S = ta->S;
      if constexpr (MATPAR) {          // form_KKT with this instance's scaled P and A (kkt.c:6-177): sigma on the x diagonal,
        const MatCtx& mc = *ta->mc;    // P (upper triangle) and A scattered to their slots of the permuted lower triangle
        const CpgMatHeader* MH = mc.mv.H;
        for (int i = lane; i < MH->n_slots; i += LANES) S[i] = __ldg(mc.mv.F64 + MH->f_S0 + i);
        __syncwarp();
        for (int e = lane; e < MH->nnzP; e += LANES) S[__ldg(mc.mv.U16 + MH->h_Pslot + e)] += mc.Pv[e];
        for (int e = lane; e < MH->nnzA; e += LANES) S[__ldg(mc.mv.U16 + MH->h_Aslot + e)] = mc.Av[e];
      } else {
        for (int i = lane; i < tv.H->n_slots; i += LANES) S[i] = tv.F64[tv.H->f_S0 + i];
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < NZL; ++k) {
        const int j = lane + 32 * k;
        if (j < M) S[tv.U16[tv.H->h_rho_slot + j]] = -rinv_of(k);
      }
      __syncwarp();
      tail_factor(tv, S, lane);
    }
  };
  if (TAIL && !MATPAR) {
    const double* ts = ta->state;
#pragma unroll
    for (int k = 0; k < NXL; ++k) { const int i = lane + 32 * k; if (i < N) s.x[k] = ts[i]; }
#pragma unroll
    for (int k = 0; k < NZL; ++k) { const int j = lane + 32 * k; if (j < M) { s.z[k] = ts[N + j]; s.y[k] = ts[N + M + j]; } }
    rho_in = ts[N + 2 * M]; rho_eq = RHO_EQ_FACTOR * rho_in; rinv_in = 1.0 / rho_in; rinv_eq = 1.0 / rho_eq;
    it0 = (int)ts[N + 2 * M + 1];
    refactor();
  } else if (st.warm_start && io.x0 != nullptr && io.y0 != nullptr) {
#pragma unroll
    for (int k = 0; k < NXL; ++k) {
      const int i = lane + 32 * k;
      if (i < N) { s.x[k] = Dinv[i] * io.x0[(size_t)b * N + i]; w[i] = s.x[k]; }
    }
#pragma unroll
    for (int k = 0; k < NZL; ++k) {
      const int j = lane + 32 * k;
      if (j < M) s.y[k] = (Einv[j] * io.y0[(size_t)b * M + j]) * c;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NZL; ++k) {
      const int j = lane + 32 * k;
      if (j < M) s.z[k] = dotA(k, w);
    }
    __syncwarp();
  }
  if constexpr (MATPAR) refactor();    // every instance owns its factor (update_matrices + update_rho_vec)

  int status = ST_UNSOLVED;
  int it = 0;
  double rho_new = rho_in;
  bool handoff = !TAIL && type_mismatch;     // (MODE 0 only)     // a constraint changed type: this instance needs its own KKT factor

  // ---- residuals + norms of the current iterate (update_info, auxil.c:564-629)
  auto update_info = [&]() {
#pragma unroll
    for (int k = 0; k < NXL; ++k) { const int i = lane + 32 * k; if (i < N) w[i] = s.x[k]; }
#pragma unroll
    for (int k = 0; k < NZL; ++k) { const int j = lane + 32 * k; if (j < M) w[N + j] = s.y[k]; }
    __syncwarp();
    double m_rp = 0, m_z = 0, m_Ax = 0, ms_rp = 0, ms_z = 0, ms_Ax = 0;
#pragma unroll
    for (int k = 0; k < NZL; ++k) {
      const int j = lane + 32 * k;
      if (j < M) {
        const double Ax = dotA(k, w);
        const double rp = Ax - s.z[k];
        const double e = unscale ? Einv[j] : 1.0;
        m_rp = fmax(m_rp, fabs(e * rp)); m_z = fmax(m_z, fabs(e * s.z[k])); m_Ax = fmax(m_Ax, fabs(e * Ax));
        ms_rp = fmax(ms_rp, fabs(rp)); ms_z = fmax(ms_z, fabs(s.z[k])); ms_Ax = fmax(ms_Ax, fabs(Ax));
      }
    }
    double m_rd = 0, m_q = 0, m_Aty = 0, m_Px = 0, ms_rd = 0, ms_q = 0, ms_Aty = 0, ms_Px = 0, a_xPx = 0, a_qx = 0;
#pragma unroll
    for (int k = 0; k < NXL; ++k) {
      const int i = lane + 32 * k;
      if (i < N) {
        const double Px = dotP(k, w);
        const double Aty = (M > 0) ? dotAt(k, w) : 0.0;
        const double rd = s.q[k] + Px + Aty;
        const double d = unscale ? Dinv[i] : 1.0;
        m_rd = fmax(m_rd, fabs(d * rd)); m_q = fmax(m_q, fabs(d * s.q[k]));
        m_Aty = fmax(m_Aty, fabs(d * Aty)); m_Px = fmax(m_Px, fabs(d * Px));
        ms_rd = fmax(ms_rd, fabs(rd)); ms_q = fmax(ms_q, fabs(s.q[k]));
        ms_Aty = fmax(ms_Aty, fabs(Aty)); ms_Px = fmax(ms_Px, fabs(Px));
        a_xPx = fma(s.x[k], Px, a_xPx); a_qx = fma(s.q[k], s.x[k], a_qx);
      }
    }
    __syncwarp();
    const double cs = unscale ? cinv : 1.0;
    s.pri_res = (M > 0) ? warp_max(m_rp) : 0.0;
    s.nrm_z = warp_max(m_z); s.nrm_Ax = warp_max(m_Ax);
    s.dua_res = cs * warp_max(m_rd);
    s.nrm_q = warp_max(m_q); s.nrm_Aty = warp_max(m_Aty); s.nrm_Px = warp_max(m_Px);
    s.s_rp = warp_max(ms_rp); s.s_z = warp_max(ms_z); s.s_Ax = warp_max(ms_Ax);
    s.s_rd = warp_max(ms_rd); s.s_q = warp_max(ms_q); s.s_Aty = warp_max(ms_Aty); s.s_Px = warp_max(ms_Px);
    s.xPx = warp_sum(a_xPx); s.qx = warp_sum(a_qx);
  };

  // ---- check_termination (auxil.c:681-786); returns the new status (ST_UNSOLVED = continue)
  auto check_termination = [&](bool approximate) -> int {
    double ea = st.eps_abs, er = st.eps_rel, epi = st.eps_prim_inf, edi = st.eps_dual_inf;
    if (approximate) { ea *= 10; er *= 10; epi *= 10; edi *= 10; }
    if (s.pri_res > OSQP_INFTY || s.dua_res > OSQP_INFTY) return ST_NONCVX;
    const double cs = unscale ? cinv : 1.0;
    bool prim_ok, prim_inf = false, dual_inf = false;
    if (M == 0) prim_ok = true;
    else {
      const double eps_prim = ea + er * fmax(s.nrm_z, s.nrm_Ax);
      prim_ok = s.pri_res < eps_prim;
      if (!prim_ok) {               // is_primal_infeasible, auxil.c:361-424
        double dproj[I::NZLs];
        double nd = 0, lhs = 0;
#pragma unroll
        for (int k = 0; k < NZL; ++k) {
          const int j = lane + 32 * k;
          double d = s.dy[k];
          const bool up_inf = s.u[k] > OSQP_INFTY * MIN_SCALING, lo_inf = s.l[k] < -OSQP_INFTY * MIN_SCALING;
          if (up_inf) d = lo_inf ? 0.0 : fmin(d, 0.0);
          else if (lo_inf) d = fmax(d, 0.0);
          if (j >= M) d = 0.0;
          dproj[k] = d;
          nd = fmax(nd, fabs((unscale && j < M) ? Ev[j] * d : d));
          lhs += s.u[k] * fmax(d, 0.0) + s.l[k] * fmin(d, 0.0);
        }
        nd = warp_max(nd);
        if (nd > DIVISION_TOL) {
          lhs = warp_sum(lhs);
          if (lhs < epi * nd) {
#pragma unroll
            for (int k = 0; k < NZL; ++k) { const int j = lane + 32 * k; if (j < M) w[N + j] = dproj[k]; }
            __syncwarp();
            double mx = 0;
#pragma unroll
            for (int k = 0; k < NXL; ++k) {
              const int i = lane + 32 * k;
              if (i < N) {
                double v = dotAt(k, w);
                if (unscale) v *= Dinv[i];
                mx = fmax(mx, fabs(v));
              }
            }
            __syncwarp();
            prim_inf = warp_max(mx) < epi * nd;
          }
        }
      }
    }
    const double eps_dual = ea + er * cs * fmax(fmax(s.nrm_q, s.nrm_Aty), s.nrm_Px);
    const bool dual_ok = s.dua_res < eps_dual;
    if (!dual_ok) {                 // is_dual_infeasible, auxil.c:426-512
      double nd = 0, qd = 0;
#pragma unroll
      for (int k = 0; k < NXL; ++k) {
        const int i = lane + 32 * k;
        if (i < N) { nd = fmax(nd, fabs(unscale ? Dv[i] * s.dx[k] : s.dx[k])); qd = fma(s.q[k], s.dx[k], qd); }
      }
      nd = warp_max(nd);
      const double cost_scaling = unscale ? c : 1.0;
      if (nd > DIVISION_TOL) {
        qd = warp_sum(qd);
        if (qd < cost_scaling * edi * nd) {
#pragma unroll
          for (int k = 0; k < NXL; ++k) { const int i = lane + 32 * k; if (i < N) w[i] = s.dx[k]; }
          __syncwarp();
          double mx = 0;
#pragma unroll
          for (int k = 0; k < NXL; ++k) {
            const int i = lane + 32 * k;
            if (i < N) {
              double v = dotP(k, w);
              if (unscale) v *= Dinv[i];
              mx = fmax(mx, fabs(v));
            }
          }
          if (warp_max(mx) < cost_scaling * edi * nd) {
            bool bad = false;
#pragma unroll
            for (int k = 0; k < NZL; ++k) {
              const int j = lane + 32 * k;
              if (j < M) {
                double v = dotA(k, w);
                if (unscale) v *= Einv[j];
                bad |= ((s.u[k] < OSQP_INFTY * MIN_SCALING) && (v > edi * nd)) ||
                       ((s.l[k] > -OSQP_INFTY * MIN_SCALING) && (v < -edi * nd));
              }
            }
            dual_inf = !__any_sync(FULL, bad);
          }
          __syncwarp();
        }
      }
    }
    if (prim_ok && dual_ok) return approximate ? ST_SOLVED_INACC : ST_SOLVED;
    if (prim_inf) return approximate ? ST_PINF_INACC : ST_PINF;
    if (dual_inf) return approximate ? ST_DINF_INACC : ST_DINF;
    return ST_UNSOLVED;
  };

  // ---- main ADMM loop (osqp.c:354-527)
  if (!handoff) {
    for (it = it0 + 1; it <= st.max_iter; ++it) {
      // compute_rhs (auxil.c:161-175), written straight into pivot order
#pragma unroll
      for (int k = 0; k < NXL; ++k) { const int i = lane + 32 * k; if (i < N) w[px[k]] = sigma * s.x[k] - s.q[k]; }
#pragma unroll
      for (int k = 0; k < NZL; ++k) { const int j = lane + 32 * k; if (j < M) w[pz[k]] = s.z[k] - rinv_of(k) * s.y[k]; }
      __syncwarp();
      tail_solve(ta->tv, ta->S, w, lane);
      // update_x, update_z (+project), update_y (auxil.c:185-225)
#pragma unroll
      for (int k = 0; k < NXL; ++k) {
        const int i = lane + 32 * k;
        if (i < N) {
          const double xn = alpha * w[px[k]] + (1.0 - alpha) * s.x[k];
          s.dx[k] = xn - s.x[k];
          s.x[k] = xn;
        }
      }
#pragma unroll
      for (int k = 0; k < NZL; ++k) {
        const int j = lane + 32 * k;
        if (j < M) {
          const double ri = rinv_of(k);
          const double zt = (s.z[k] - ri * s.y[k]) + ri * w[pz[k]];     // z~ = rhs_z + rho^-1 nu (qdldl_interface.c:368-370)
          const double v = alpha * zt + (1.0 - alpha) * s.z[k];
          const double zn = fmin(fmax(v + ri * s.y[k], s.l[k]), s.u[k]);
          const double r = ((s.loosemask >> k) & 1u) ? RHO_MIN : (((s.eqmask >> k) & 1u) ? rho_eq : rho_in);
          s.dy[k] = r * (v - zn);
          s.y[k] += s.dy[k];
          s.z[k] = zn;
        }
      }
      __syncwarp();
      const bool can_check = st.check_termination && (it % st.check_termination == 0);
      const bool can_adapt = st.adaptive_rho && st.adaptive_rho_interval && (it % st.adaptive_rho_interval == 0);
      if (can_check || can_adapt || it == st.max_iter) {
        update_info();
        if (can_check || it == st.max_iter) {
          status = check_termination(false);
          if (status != ST_UNSOLVED) break;
        }
        if (can_adapt) {            // compute_rho_estimate + adapt_rho decision (auxil.c:13-74)
          const double pn = s.s_rp / (fmax(s.s_z, s.s_Ax) + DIVISION_TOL);
          const double dn = s.s_rd / (fmax(fmax(s.s_q, s.s_Aty), s.s_Px) + DIVISION_TOL);
          double r = rho_in * sqrt(pn / dn);
          r = fmin(fmax(r, RHO_MIN), RHO_MAX);
          if (r > rho_in * st.adaptive_rho_tolerance || r < rho_in / st.adaptive_rho_tolerance) {
            if (TAIL) {                              // osqp_update_rho (osqp.c:1268-1325) + refactor
              rho_in = r; rho_eq = RHO_EQ_FACTOR * r; rinv_in = 1.0 / rho_in; rinv_eq = 1.0 / rho_eq;
              refactor();
            } else {
              rho_new = r; handoff = true; break;   // needs a per-instance refactorisation: tail kernel
            }
          }
        }
      }
    }
    if (it > st.max_iter) it = st.max_iter;
    if (!handoff && status == ST_UNSOLVED) {      // osqp.c:563-568
      status = check_termination(true);
      if (status == ST_UNSOLVED) status = ST_MAXITER;
    }
  }

  if (handoff) {
    int slot = -1;
    if (lane == 0) slot = atomicAdd(io.tail_count, 1);
    slot = __shfl_sync(FULL, slot, 0);
    if (lane == 0) { io.status[b] = ST_HANDOFF; io.iter[b] = it; }
    if (slot < io.tail_capacity) {
      double* ts = io.tail_state + (size_t)slot * (N + 2 * M + 2);
#pragma unroll
      for (int k = 0; k < NXL; ++k) { const int i = lane + 32 * k; if (i < N) ts[i] = s.x[k]; }
#pragma unroll
      for (int k = 0; k < NZL; ++k) { const int j = lane + 32 * k; if (j < M) { ts[N + j] = s.z[k]; ts[N + M + j] = s.y[k]; } }
      if (lane == 0) { ts[N + 2 * M] = rho_new; ts[N + 2 * M + 1] = (double)it; io.tail_ids[slot] = b; }
    }
    return;
  }

  // ---- store_solution / unscale_solution (auxil.c:524-562, scaling.c:177-192) + retrieval (a11, a12)
  const bool has_sol = !(status == ST_PINF || status == ST_PINF_INACC || status == ST_DINF ||
                         status == ST_DINF_INACC || status == ST_NONCVX);
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
  for (int k = 0; k < NXL; ++k) {
    const int i = lane + 32 * k;
    if (i < N) {
      const double xv = has_sol ? Dv[i] * s.x[k] : qnan;
      w[i] = xv;
      if (io.sol_x) io.sol_x[(size_t)b * N + i] = xv;
    }
  }
#pragma unroll
  for (int k = 0; k < NZL; ++k) {
    const int j = lane + 32 * k;
    if (j < M) {
      const double yv = has_sol ? (Ev[j] * s.y[k]) * cinv : qnan;
      w[N + j] = yv;
      if (io.sol_y) io.sol_y[(size_t)b * M + j] = yv;
    }
  }
  __syncwarp();
  if (io.prim) {
    const int np = H->n_prim;
    for (int k = lane; k < np; k += LANES) io.prim[(size_t)b * np + k] = w[U16[H->h_prim + k]];
  }
  if (io.dual) {
    const int nd = H->n_dual;
    for (int k = lane; k < nd; k += LANES) io.dual[(size_t)b * nd + k] = w[N + U16[H->h_dual + k]];
  }
  __syncwarp();
  if (lane == 0) {
    double obj = (0.5 * s.xPx + s.qx);                          // compute_obj_val, auxil.c:227-238
    if (st.scaling) obj *= cinv;
    if (status == ST_PINF || status == ST_PINF_INACC) obj = OSQP_INFTY;
    else if (status == ST_DINF || status == ST_DINF_INACC) obj = -OSQP_INFTY;
    else if (status == ST_NONCVX) obj = qnan;
    else obj = (H->is_max ? -1.0 : 1.0) * (obj + H->d_const);   // cpg_retrieve_info, cvxpygen/utils.py:980
    io.obj_val[b] = obj; io.iter[b] = it; io.status[b] = status;
    io.pri_res[b] = s.pri_res; io.dua_res[b] = s.dua_res;
  }
}

// BIG families: every instance goes to the per-instance-factor kernel from iteration 0 (cold start, the family's rho)
__global__ void queue_all_kernel(const BatchIO io, int words, int rho_word, double rho) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= io.B) return;
  double* ts = io.tail_state + (size_t)b * words;
  for (int k = 0; k < words; ++k) ts[k] = 0.0;
  ts[rho_word] = rho;
  io.tail_ids[b] = b;
  io.status[b] = ST_HANDOFF; io.iter[b] = 0;
  if (b == 0) *io.tail_count = io.B;
}

// Tail kernel: instances whose rho changed (or whose bounds changed a constraint type) continue here with their
// own numeric factor.  One warp per instance; reads the hand-off queue written by admm_batch_kernel.
template <class Fam>
__global__ void __launch_bounds__(Fam::TAIL_WARPS * 32, 1)
admm_tail_kernel(const uint8_t* __restrict__ blob_g, const uint8_t* __restrict__ tail_blob_g,
                 const BatchIO io, const Settings st) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const int n_tail = min(*io.tail_count, io.tail_capacity);
  if (n_tail == 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint8_t* base = blob_g;
  int ws_off = 0;
  if constexpr (Fam::TAIL_STAGE) {     // the compact constants blob fits next to the warps' workspaces: stage it with TMA
    const uint32_t total = reinterpret_cast<const CpgBlobHeader*>(blob_g)->total_bytes;
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(&bar, total);
      constexpr uint32_t CHUNK = 32768;
      for (uint32_t off = 0; off < total; off += CHUNK)
        tma_bulk_g2s(smem + off, blob_g + off, (total - off < CHUNK) ? (total - off) : CHUNK, &bar);
    }
    mbar_wait(&bar, 0);
    base = smem; ws_off = Fam::CBLOB_BYTES_PAD;
  }                                    // else (large families): the tables are read through L1 / L2 where they lie
  const CpgBlobHeader* H = reinterpret_cast<const CpgBlobHeader*>(base);
  const int* I32 = reinterpret_cast<const int*>(base + H->off_i32);
  const double* F64 = reinterpret_cast<const double*>(base + H->off_f64);
  const uint16_t* U16 = reinterpret_cast<const uint16_t*>(base + H->off_u16);
  double* wbase = reinterpret_cast<double*>(smem + ws_off) + (size_t)warp * (Fam::W_STRIDE + Fam::S_STRIDE);
  TailArgs ta;
  ta.tv = make_tail_view(tail_blob_g);
  ta.S = wbase + Fam::W_STRIDE;
  ta.mc = nullptr;
  for (int slot = blockIdx.x * Fam::TAIL_WARPS + warp; slot < n_tail; slot += gridDim.x * Fam::TAIL_WARPS) {
    ta.state = io.tail_state + (size_t)slot * (Fam::N + 2 * Fam::M + 2);
    const int b = io.tail_ids[slot];
    solve_instance<Fam, 1>(H, I32, F64, U16, wbase, lane, b, io, st, &ta);
    __syncwarp();
  }
}

}  // namespace cpgb200
