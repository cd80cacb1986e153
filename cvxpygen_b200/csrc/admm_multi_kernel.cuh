// admm_multi_kernel.cuh -- main batched ADMM kernel for sm_100a: NI problem instances per warp (NI = 2 or 4).
//
// Why several instances per warp: ncu on the one-instance-per-warp version showed the kernel bound by shared-memory
// bandwidth (71 % of peak wavefronts; profiles/r1_v1_ncu_summary.md): every inner step of the KKT-solve schedule loads
// one 8-byte coefficient per lane and uses it once.  Register-blocking NI instances makes every coefficient (and index)
// loaded from shared memory feed NI FMAs; the instances' work vectors are interleaved (wN[pos * NI + s]) so that one or
// two LDS.128 fetch all operands.  The chain part of the schedule is encoded as DENSE tiles whose lanes read the same
// wN address (a shared-memory broadcast) and load no index at all; the whole schedule is emitted as straight-line
// code with immediate offsets (cpg_kkt_solve_gen.cuh, offline/emit_solve.py).
//
// The NI slots of a warp are independent instances: each has its own iteration counter, termination check and
// epilogue; when one terminates its slot is refilled from the global instance queue at once.
// State per slot in registers: x and the pre-projection vector t, of which z and y are functions (lane i%32 owns element i).  q, l, u are NOT kept in registers: rows that do
// not depend on a batched parameter are read from the constants blob (already scaled), rows that do are read from a
// small per-warp table written by the slot's prologue.
// Everything that happens once per instance or once per check_termination iterations (prologue, residuals,
// termination + infeasibility tests, hand-off, epilogue) is written for SLOT 0 only; the slots are brought to
// position 0 one after the other by rotating the register state.  That keeps the cold code NI times smaller (the hot
// loop shares the instruction cache with it) at the cost of ~100 register moves per rotation.
//
// Reference functions restated: same list as admm_kernel.cuh (a1-a12); the per-instance arithmetic and its order are
// those of the single-instance path except that z, y are recomputed from t (agreement to ~1e-14, identical iteration counts).
#pragma once
#include <type_traits>
#include "admm_kernel.cuh"

namespace cpgb200 {

template <int NI> struct MultiOps;
template <> struct MultiOps<2> {
  static __device__ __forceinline__ void ld(const double* p, double (&o)[2]) {
    const double2 a = *reinterpret_cast<const double2*>(p); o[0] = a.x; o[1] = a.y;
  }
  static __device__ __forceinline__ void st(double* p, const double (&v)[2]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
  }
};
template <> struct MultiOps<4> {
  static __device__ __forceinline__ void ld(const double* p, double (&o)[4]) {
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  }
  static __device__ __forceinline__ void st(double* p, const double (&v)[4]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
  }
};

}  // namespace cpgb200
#ifndef CPG_FAM_BIG
#define CPG_FAM_BIG 0
#endif
#ifndef CPG_FAM_DMMA
#define CPG_FAM_DMMA 0
#endif
// the generated straight-line schedule is compiled only for the kernel that runs it: not for the tensor-core variant and not for
// BIG families (schedule larger than shared memory: solved by the per-instance-factor kernel alone, nvcc never sees their
// hundreds of thousands of generated lines)
#define CPG_FAM_STRAIGHT (!CPG_FAM_DMMA && !CPG_FAM_BIG)
#if CPG_FAM_STRAIGHT
#include "cpg_kkt_solve_gen.cuh"     // straight-line schedule of this family, template <int NI>
#endif
namespace cpgb200 {

// ================================================================ FP64 tensor-core KKT solve (CPG_FAM_DMMA families)
// A GROUP of four warps owns eight instances; their work vectors are interleaved, w8[position][instance], so that the
// operand fragment of mma.sync.m8n8k4.f64 (4 positions x 8 instances) is four 64-byte rows and the 8x8 result fragment is
// eight.  Schedule and table formats: offline/dmma.py.  The 16-byte pair of instances (2c, 2c+1) of position p sits at
// column 2c ^ 2*((p >> 2) & 3): the ADMM phases, where a warp touches ONE pair of 32 scattered positions, then spread over
// all banks instead of four, and the MMA fragments stay conflict-free (the XOR permutes pairs inside one 64-byte row).
struct DmmaCtx {
  double* w8;                 // this group's interleaved work vectors
  double* stage;              // this group's staging buffer: 8x8 partial results, fragment order
  const uint4* items;         // [stream index][warp of group] -> {mask, value offset, pos0|pos1<<16, pos2|pos3<<16}
  const double* vals;         // compressed coefficients
  const int4* tile_hdr;       // {item base, rounds | job rounds << 8, round_len base, job base}
  const uint16_t* round_len;
  const uint32_t* jobs;       // [job round][warp of group][8 words]
  int* gflag;                 // this group's two flag words (bit 0: a warp checks termination this iteration, bit 1: alive)
  int n_tiles, wg, bar_id;
};

__device__ __forceinline__ int w8_off(int p, int c) { return p * 8 + (c ^ (((p >> 2) & 3) << 1)); }

#ifdef CPG_SIMT_HOST_EMU
__device__ inline void group_bar(int id) { simt::named_barrier(id, 128); }
// D(8x8) += A(8x4) B(4x8): lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2(l%4) + {0,1}]  (PTX ISA, mma.m8n8k4 .f64 fragments)
__device__ inline void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  const int lane = threadIdx.x & 31, row = lane >> 2, kk = lane & 3;
  for (int j = 0; j < 4; ++j) {
    const double aj = __shfl_sync(FULL, a, row * 4 + j);
    const double b0 = __shfl_sync(FULL, b, (2 * kk) * 4 + j), b1 = __shfl_sync(FULL, b, (2 * kk + 1) * 4 + j);
    d0 = fma(aj, b0, d0); d1 = fma(aj, b1, d1);
  }
}
#else
__device__ __forceinline__ void group_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#endif

// Shared-memory accesses of the solve by 32-bit shared-window address: the function is not inlined (its item loop wants its own
// register allocation -- inlined, the 72 state registers of the ADMM phases squeeze it into rematerialising lane constants), and
// through generic pointers every access would be a 64-bit LD.E with carry chains.
#ifdef CPG_SIMT_HOST_EMU
typedef uintptr_t saddr_t;
__device__ inline saddr_t s_addr(const void* p) { return reinterpret_cast<uintptr_t>(p); }
__device__ inline uint4 lds_u4(saddr_t a) { return *reinterpret_cast<const uint4*>(a); }
__device__ inline int4 lds_i4(saddr_t a) { return *reinterpret_cast<const int4*>(a); }
__device__ inline unsigned lds_u32(saddr_t a) { return *reinterpret_cast<const unsigned*>(a); }
__device__ inline unsigned lds_u16(saddr_t a) { return *reinterpret_cast<const uint16_t*>(a); }
__device__ inline double lds_f64(saddr_t a) { return *reinterpret_cast<const double*>(a); }
__device__ inline double2 lds_f64x2(saddr_t a) { return *reinterpret_cast<const double2*>(a); }
__device__ inline void sts_f64x2(saddr_t a, double x, double y) { *reinterpret_cast<double2*>(a) = make_double2(x, y); }
#else
typedef uint32_t saddr_t;
__device__ __forceinline__ saddr_t s_addr(const void* p) { return (saddr_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 lds_u4(saddr_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ int4 lds_i4(saddr_t a) { int4 v; asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ unsigned lds_u32(saddr_t a) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ unsigned lds_u16(saddr_t a) { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return v; }
__device__ __forceinline__ double lds_f64(saddr_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ double2 lds_f64x2(saddr_t a) { double2 v; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void sts_f64x2(saddr_t a, double x, double y) { asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(x), "d"(y) : "memory"); }
#endif

#ifndef DMMA_SOLVE_INLINE
#define DMMA_SOLVE_INLINE __noinline__
#endif
// one KKT solve for the group's eight right-hand sides, in place in w8; every warp of the group calls it; ends on a group barrier.
// The loop is instruction-issue bound (profiles/r2_dmma_v1_ncu_summary.md), so an item costs as few instructions as the table
// format allows: one broadcast LDS.128 (descriptor), PRMT + 2 LOP3 + IADD + LDS.64 (operand: the descriptor carries the swizzled
// byte offsets of its four rows, the lane XORs its instance column in), LOP3 + POPC + shift/add + LDS.64 + 2 FSEL (compressed
// coefficient: rank of the lane's bit in the occupancy mask; loaded unconditionally, then selected), DMMA.
__device__ DMMA_SOLVE_INLINE void dmma_solve(const DmmaCtx& dc, const int lane) {
  const int kk = lane & 3, nn = lane >> 2;
  const unsigned lt = (1u << lane) - 1u, lbit = 1u << lane;
  const unsigned sel = 0x1010u + 0x2222u * kk;            // byte selector: 16-bit field kk of the pair of words (z, w)
  const unsigned ncol = (unsigned)nn << 3;
  const saddr_t vals = s_addr(dc.vals), w8b = s_addr(dc.w8);
  const saddr_t stage = s_addr(dc.stage) + lane * 16;
  const saddr_t hdr = s_addr(dc.tile_hdr), rl = s_addr(dc.round_len), jobs = s_addr(dc.jobs), items = s_addr(dc.items);
  const int wg = dc.wg, bar_id = dc.bar_id, n_tiles = dc.n_tiles;
  for (int t = 0; t < n_tiles; ++t) {
    const int4 th = lds_i4(hdr + t * 16);
    const int n_rounds = th.y & 0xff, n_jr = th.y >> 8;
    saddr_t it = items + (th.x * 4 + wg) * 16;
    for (int r = 0; r < n_rounds; ++r) {
      const int L = (int)lds_u16(rl + (th.z + r) * 2);
      double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;        // two accumulator chains
#pragma unroll 2
      for (int i = 0; i < L; i += 2) {
        const uint4 d0 = lds_u4(it), d1 = lds_u4(it + 64);
        it += 128;
        const double b0 = lds_f64(w8b + ((__byte_perm(d0.z, d0.w, sel) & 0xffffu) ^ ncol));
        const double b1 = lds_f64(w8b + ((__byte_perm(d1.z, d1.w, sel) & 0xffffu) ^ ncol));
        // the load is unconditional (a lane without a coefficient reads a neighbour's -- always inside the table) and the
        // result is selected: a predicated load costs a divergent branch here
        double a0 = lds_f64(vals + d0.y + 8 * __popc(d0.x & lt));
        double a1 = lds_f64(vals + d1.y + 8 * __popc(d1.x & lt));
        a0 = (d0.x & lbit) ? a0 : 0.0;
        a1 = (d1.x & lbit) ? a1 : 0.0;
        dmma_8x8x4(c00, c01, a0, b0);
        dmma_8x8x4(c10, c11, a1, b1);
      }
      sts_f64x2(stage + (wg * n_rounds + r) * 512, c00 + c10, c01 + c11);
    }
    group_bar(bar_id);                    // every operand of the tile has been read, every partial result is staged
    saddr_t jb = jobs + (th.w * 4 + wg) * 32;
    for (int j = 0; j < n_jr; ++j, jb += 128) {
      const unsigned row = __byte_perm(lds_u32(jb + (nn >> 1) * 4), 0u, (nn & 1) ? 0x4432u : 0x4410u);   // swizzled byte offset of result row lane/4
      const unsigned sl = lds_u32(jb + 16);
      const int parts = (int)lds_u32(jb + 20);
      double2 acc = lds_f64x2(stage + (sl & 0xffu) * 512);
      if (parts > 1) {
        const double2 v = lds_f64x2(stage + ((sl >> 8) & 0xffu) * 512);
        acc.x += v.x; acc.y += v.y;
        if (parts > 2) {
          const double2 v2 = lds_f64x2(stage + ((sl >> 16) & 0xffu) * 512);
          const double2 u2 = lds_f64x2(stage + (sl >> 24) * 512);
          acc.x += v2.x; acc.y += v2.y; acc.x += u2.x; acc.y += u2.y;
        }
      }
      if (row != 0xffffu) sts_f64x2(w8b + (row ^ ((unsigned)kk << 4)), acc.x, acc.y);
    }
    group_bar(bar_id);                    // the rows of this tile are in place before the next tile reads them
  }
}

// row-blocked ELL sparse mat-vec on an interleaved multi-vector; returns the result for column 0 only
template <int NI>
__device__ __forceinline__ double ell_dot_col0(const int* tab, const double* F64, const uint16_t* U16,
                                               const double* vecN, int lane) {
  const int K = tab[0];
  const double* v = F64 + tab[1] + lane;
  const uint16_t* c = U16 + tab[2] + lane;
  double a = 0.0;
  for (int k = 0; k < K; ++k) a = fma(v[k * LANES], vecN[NI * c[k * LANES]], a);
  return a;
}

// DMMA = false: every warp solves for its NI instances on its own (generated straight-line schedule, wN = its work vectors).
// DMMA = true : NI = 2; the four warps of a group solve their eight right-hand sides together (dmma_solve); wN is then the
//               warp's PRIVATE scratch for the once-per-check routines (a quarter of the group's w8, free between solves).
template <class Fam, int NI, bool DMMA = false>
__device__ void solve_multi(const CpgBlobHeader* __restrict__ H, const int* __restrict__ I32,
                            const double* __restrict__ F64, const uint16_t* __restrict__ U16,
                            double* __restrict__ wN, double* __restrict__ bv, const int lane,
                            const BatchIO& io, const Settings& st, const DmmaCtx* dc = nullptr) {
  static_assert(!DMMA || NI == 2, "the tensor-core path pairs two instances per warp, eight per group");
  constexpr int N = Fam::N, M = Fam::M, NXL = (N + 31) / 32, NZL = (M + 31) / 32, NZLs = NZL > 0 ? NZL : 1;
  using MO = MultiOps<NI>;
  const double* Dv = F64 + H->f_D;  const double* Dinv = F64 + H->f_Dinv;
  const double* Ev = F64 + H->f_E;  const double* Einv = F64 + H->f_Einv;
  const double c = H->c, cinv = H->cinv, sigma = H->sigma, alpha = st.alpha;
  const bool unscale = st.scaling && !st.scaled_termination;
  const double rho_in = H->rho, rho_eq = RHO_EQ_FACTOR * H->rho;
  const double rinv_in = 1.0 / rho_in, rinv_eq = 1.0 / rho_eq, rinv_loose = 1.0 / RHO_MIN;

  // ---- per-lane constants of the family (identical for every instance), read from the blob when needed
  // addr tables: >= 0 index into the F64 area (row constant over the batch); < 0 -(slot+1) in the per-warp table bv.
  const uint16_t* PX = U16 + H->h_pinvx + lane;
  const uint16_t* PZ = U16 + H->h_pinvz + lane;
  const int* AQ = I32 + H->i_addr_q + lane;
  const int* AL = I32 + H->i_addr_l + lane;
  const int* AU = I32 + H->i_addr_u + lane;
  unsigned eqmask = 0u, loosemask = 0u;
#pragma unroll
  for (int k = 0; k < NZL; ++k) {
    const int j = lane + 32 * k;
    if (j < M) {
      const int ct = U16[H->h_ctype + j];
      if (ct == 2) eqmask |= 1u << k;
      if (ct == 0) loosemask |= 1u << k;
    }
  }
  auto tabN = [&](const int* T, int k, double (&o)[NI]) __attribute__((always_inline)) {
    const int a = T[32 * k];
    if (a >= 0) {
      const double v = F64[a];
#pragma unroll
      for (int s = 0; s < NI; ++s) o[s] = v;
    } else MO::ld(bv + NI * (-a - 1), o);
  };
  auto tab0 = [&](const int* T, int k) __attribute__((always_inline)) -> double {       // slot 0 only
    const int a = T[32 * k];
    return (a >= 0) ? F64[a] : bv[NI * (-a - 1)];
  };
  auto rinv_of = [&](int k) __attribute__((always_inline)) -> double {
    return ((loosemask >> k) & 1u) ? rinv_loose : (((eqmask >> k) & 1u) ? rinv_eq : rinv_in); };
  auto rho_of = [&](int k) __attribute__((always_inline)) -> double {
    return ((loosemask >> k) & 1u) ? RHO_MIN : (((eqmask >> k) & 1u) ? rho_eq : rho_in); };

  // ---- per-slot state: x and the PRE-PROJECTION vector t = alpha z~ + (1 - alpha) z_prev + y_prev / rho of the last update
  // (auxil.c:197-213 computes exactly this value before projecting it).  z and y are functions of it,
  //     z = clip(t, l, u),      y = rho (t - z)
  // so one vector per instance stays in registers instead of two (profile of the (x, z, y) version: 87 M spilled loads per
  // 100 000 instances).  The start point of an instance is not of that form (z0 = 0 or A x0 need not lie in [l, u]): until its
  // first update a slot is flagged `first` and its (z0, y0 / rho) are 0 (cold start) or read from the instance's scratch rows
  // io.ws (warm start).
  double x[NI][NXL], t[NI][NZLs];
  int inst[NI], it[NI];
  bool active[NI], first[NI];
  bool exhausted = false;
  const bool warm = st.warm_start && io.x0 != nullptr && io.y0 != nullptr;
#pragma unroll
  for (int s = 0; s < NI; ++s) {
    inst[s] = -1; it[s] = 0; active[s] = false; first[s] = true;
#pragma unroll
    for (int k = 0; k < NXL; ++k) x[s][k] = 0.0;
#pragma unroll
    for (int k = 0; k < NZLs; ++k) t[s][k] = 0.0;
  }
  // z and d = y / rho of slot s, row lane + 32 k, given the row's bounds
  auto zd_of = [&](int s, int k, double lv, double uv, double& zv, double& dv) __attribute__((always_inline)) {
    zv = fmin(fmax(t[s][k], lv), uv); dv = t[s][k] - zv;
  };
  auto zd_first = [&](int s, int k, double lv, double uv, double& zv, double& dv) __attribute__((always_inline)) {
    if (first[s]) {
      zv = 0.0; dv = 0.0;
      const int j = lane + 32 * k;
      if (warm && inst[s] >= 0 && j < M) { const double* wr = io.ws + (size_t)inst[s] * (2 * M); zv = wr[j]; dv = wr[M + j]; }
    } else zd_of(s, k, lv, uv, zv, dv);
  };

  // bring slot s+1 to position s (cyclically) -- registers and the per-warp batched-row table
  auto rotate_state = [&]() __attribute__((always_inline)) {
#pragma unroll
    for (int k = 0; k < NXL; ++k) { const double v = x[0][k];
#pragma unroll
      for (int s = 0; s + 1 < NI; ++s) x[s][k] = x[s + 1][k];
      x[NI - 1][k] = v; }
#pragma unroll
    for (int k = 0; k < NZLs; ++k) { const double v = t[0][k];
#pragma unroll
      for (int s = 0; s + 1 < NI; ++s) t[s][k] = t[s + 1][k];
      t[NI - 1][k] = v; }
    { const int v = inst[0], u = it[0]; const bool a = active[0], f = first[0];
#pragma unroll
      for (int s = 0; s + 1 < NI; ++s) { inst[s] = inst[s + 1]; it[s] = it[s + 1]; active[s] = active[s + 1]; first[s] = first[s + 1]; }
      inst[NI - 1] = v; it[NI - 1] = u; active[NI - 1] = a; first[NI - 1] = f; }
    const int nb = H->nb_slots;
    for (int r = lane; r < nb; r += LANES) {
      double v[NI], w[NI];
      MO::ld(bv + NI * r, v);
#pragma unroll
      for (int s = 0; s < NI; ++s) w[s] = v[(s + 1) % NI];
      MO::st(bv + NI * r, w);
    }
    __syncwarp();
  };

  // ================================================================ slot-0 routines (cold path)
  // a1-a3: canonicalise the batched rows of slot 0, scale them, detect constraint-type changes
  auto load_instance0 = [&](int b) __attribute__((always_inline)) -> bool {
    const double* th = io.params + (size_t)b * H->npb;
    const int nbq = H->n_bq, nbc = H->n_bc;
    for (int r = lane; r < nbq; r += LANES) {
      const int i = U16[H->h_bq_row + r];
      double acc = F64[H->f_qbase + i];
      for (int e = I32[H->i_bq_ptr + r]; e < I32[H->i_bq_ptr + r + 1]; ++e)
        acc = fma(F64[H->f_bq_val + e], __ldg(th + U16[H->h_bq_col + e]), acc);
      bv[NI * r] = (Dv[i] * acc) * c;
    }
    bool mismatch = false;
    for (int r = lane; r < nbc; r += LANES) {
      const int j = U16[H->h_bc_row + r];
      double al = F64[H->f_lbase + j], au = F64[H->f_ubase + j];
      for (int e = I32[H->i_bl_ptr + r]; e < I32[H->i_bl_ptr + r + 1]; ++e)
        al = fma(F64[H->f_bl_val + e], __ldg(th + U16[H->h_bl_col + e]), al);
      for (int e = I32[H->i_bu_ptr + r]; e < I32[H->i_bu_ptr + r + 1]; ++e)
        au = fma(F64[H->f_bu_val + e], __ldg(th + U16[H->h_bu_col + e]), au);
      al = Ev[j] * fmin(fmax(al, -OSQP_INFTY), OSQP_INFTY);
      au = Ev[j] * fmin(fmax(au, -OSQP_INFTY), OSQP_INFTY);
      bv[NI * (nbq + r)] = al;
      bv[NI * (nbq + nbc + r)] = au;
      const bool loose = (al < -OSQP_INFTY * MIN_SCALING) && (au > OSQP_INFTY * MIN_SCALING);
      const bool eq = !loose && (au - al < RHO_TOL);
      mismatch |= ((loose ? 0 : (eq ? 2 : 1)) != (int)U16[H->h_ctype + j]);
    }
    __syncwarp();
    return __any_sync(FULL, mismatch);
  };

  // z and y of slot 0 (start values while the slot is `first`)
  auto zy0 = [&](double (&zc)[NZLs], double (&yc)[NZLs]) __attribute__((always_inline)) {
#pragma unroll
    for (int k = 0; k < NZLs; ++k) {
      zc[k] = 0.0; yc[k] = 0.0;
      if (k < NZL && lane + 32 * k < M) {
        double dv;
        zd_first(0, k, tab0(AL, k), tab0(AU, k), zc[k], dv);
        yc[k] = rho_of(k) * dv;
      }
    }
  };

  // residuals + norms of slot 0's current iterate (update_info, auxil.c:564-629)
  double pri_res, dua_res, xPx, qx, nrm_z, nrm_Ax, nrm_q, nrm_Aty, nrm_Px, s_rp, s_rd, s_z, s_Ax, s_q, s_Aty, s_Px;
  auto update_info0 = [&](const double (&zc)[NZLs], const double (&yc)[NZLs]) __attribute__((always_inline)) {
#pragma unroll
    for (int k = 0; k < NXL; ++k) { const int i = lane + 32 * k; if (i < N) wN[NI * i] = x[0][k]; }
#pragma unroll
    for (int k = 0; k < NZL; ++k) { const int j = lane + 32 * k; if (j < M) wN[NI * (N + j)] = yc[k]; }
    __syncwarp();
    double m_rp = 0, m_z = 0, m_Ax = 0, ms_rp = 0, ms_z = 0, ms_Ax = 0;
#pragma unroll
    for (int k = 0; k < NZL; ++k) {
      const int j = lane + 32 * k;
      if (j < M) {
        const double Ax = ell_dot_col0<NI>(I32 + H->i_ellA + 3 * k, F64, U16, wN, lane);
        const double rp = Ax - zc[k];
        const double e = unscale ? Einv[j] : 1.0;
        m_rp = fmax(m_rp, fabs(e * rp)); m_z = fmax(m_z, fabs(e * zc[k])); m_Ax = fmax(m_Ax, fabs(e * Ax));
        ms_rp = fmax(ms_rp, fabs(rp)); ms_z = fmax(ms_z, fabs(zc[k])); ms_Ax = fmax(ms_Ax, fabs(Ax));
      }
    }
    double m_rd = 0, m_q = 0, m_Aty = 0, m_Px = 0, ms_rd = 0, ms_q = 0, ms_Aty = 0, ms_Px = 0, a_xPx = 0, a_qx = 0;
#pragma unroll
    for (int k = 0; k < NXL; ++k) {
      const int i = lane + 32 * k;
      if (i < N) {
        const double Px = ell_dot_col0<NI>(I32 + H->i_ellP + 3 * k, F64, U16, wN, lane);
        const double Aty = (M > 0) ? ell_dot_col0<NI>(I32 + H->i_ellAt + 3 * k, F64, U16, wN, lane) : 0.0;
        const double qv = tab0(AQ, k);
        const double rd = qv + Px + Aty;
        const double d = unscale ? Dinv[i] : 1.0;
        m_rd = fmax(m_rd, fabs(d * rd)); m_q = fmax(m_q, fabs(d * qv));
        m_Aty = fmax(m_Aty, fabs(d * Aty)); m_Px = fmax(m_Px, fabs(d * Px));
        ms_rd = fmax(ms_rd, fabs(rd)); ms_q = fmax(ms_q, fabs(qv));
        ms_Aty = fmax(ms_Aty, fabs(Aty)); ms_Px = fmax(ms_Px, fabs(Px));
        a_xPx = fma(x[0][k], Px, a_xPx); a_qx = fma(qv, x[0][k], a_qx);
      }
    }
    __syncwarp();
    const double cs = unscale ? cinv : 1.0;
    pri_res = (M > 0) ? warp_max(m_rp) : 0.0;
    nrm_z = warp_max(m_z); nrm_Ax = warp_max(m_Ax);
    dua_res = cs * warp_max(m_rd);
    nrm_q = warp_max(m_q); nrm_Aty = warp_max(m_Aty); nrm_Px = warp_max(m_Px);
    s_rp = warp_max(ms_rp); s_z = warp_max(ms_z); s_Ax = warp_max(ms_Ax);
    s_rd = warp_max(ms_rd); s_q = warp_max(ms_q); s_Aty = warp_max(ms_Aty); s_Px = warp_max(ms_Px);
    xPx = warp_sum(a_xPx); qx = warp_sum(a_qx);
  };

  // check_termination for slot 0 (auxil.c:681-786); dx0, dy0 = last step of slot 0
  auto check_termination0 = [&](const double (&dx0)[NXL], const double (&dy0)[NZLs], bool approximate)
      __attribute__((always_inline)) -> int {
    double ea = st.eps_abs, er = st.eps_rel, epi = st.eps_prim_inf, edi = st.eps_dual_inf;
    if (approximate) { ea *= 10; er *= 10; epi *= 10; edi *= 10; }
    if (pri_res > OSQP_INFTY || dua_res > OSQP_INFTY) return ST_NONCVX;
    const double cs = unscale ? cinv : 1.0;
    bool prim_ok, prim_inf = false, dual_inf = false;
    if (M == 0) prim_ok = true;
    else {
      prim_ok = pri_res < ea + er * fmax(nrm_z, nrm_Ax);
      if (!prim_ok) {               // is_primal_infeasible, auxil.c:361-424
        double dproj[NZLs];
        double nd = 0, lhs = 0;
#pragma unroll
        for (int k = 0; k < NZL; ++k) {
          const int j = lane + 32 * k;
          double d = 0.0;
          if (j < M) {
            d = dy0[k];
            const double lv = tab0(AL, k), uv = tab0(AU, k);
            const bool up_inf = uv > OSQP_INFTY * MIN_SCALING, lo_inf = lv < -OSQP_INFTY * MIN_SCALING;
            if (up_inf) d = lo_inf ? 0.0 : fmin(d, 0.0);
            else if (lo_inf) d = fmax(d, 0.0);
            nd = fmax(nd, fabs(unscale ? Ev[j] * d : d));
            lhs += uv * fmax(d, 0.0) + lv * fmin(d, 0.0);
          }
          dproj[k] = d;
        }
        nd = warp_max(nd);
        if (nd > DIVISION_TOL) {
          lhs = warp_sum(lhs);
          if (lhs < epi * nd) {
#pragma unroll
            for (int k = 0; k < NZL; ++k) { const int j = lane + 32 * k; if (j < M) wN[NI * (N + j)] = dproj[k]; }
            __syncwarp();
            double mx = 0;
#pragma unroll
            for (int k = 0; k < NXL; ++k) {
              const int i = lane + 32 * k;
              if (i < N) {
                double v = ell_dot_col0<NI>(I32 + H->i_ellAt + 3 * k, F64, U16, wN, lane);
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
    const bool dual_ok = dua_res < ea + er * cs * fmax(fmax(nrm_q, nrm_Aty), nrm_Px);
    if (!dual_ok) {                 // is_dual_infeasible, auxil.c:426-512
      double nd = 0, qd = 0;
#pragma unroll
      for (int k = 0; k < NXL; ++k) {
        const int i = lane + 32 * k;
        if (i < N) { nd = fmax(nd, fabs(unscale ? Dv[i] * dx0[k] : dx0[k])); qd = fma(tab0(AQ, k), dx0[k], qd); }
      }
      nd = warp_max(nd);
      const double cost_scaling = unscale ? c : 1.0;
      if (nd > DIVISION_TOL) {
        qd = warp_sum(qd);
        if (qd < cost_scaling * edi * nd) {
#pragma unroll
          for (int k = 0; k < NXL; ++k) { const int i = lane + 32 * k; if (i < N) wN[NI * i] = dx0[k]; }
          __syncwarp();
          double mx = 0;
#pragma unroll
          for (int k = 0; k < NXL; ++k) {
            const int i = lane + 32 * k;
            if (i < N) {
              double v = ell_dot_col0<NI>(I32 + H->i_ellP + 3 * k, F64, U16, wN, lane);
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
                double v = ell_dot_col0<NI>(I32 + H->i_ellA + 3 * k, F64, U16, wN, lane);
                if (unscale) v *= Einv[j];
                bad |= ((tab0(AU, k) < OSQP_INFTY * MIN_SCALING) && (v > edi * nd)) ||
                       ((tab0(AL, k) > -OSQP_INFTY * MIN_SCALING) && (v < -edi * nd));
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

  // hand slot 0 to the tail kernel (own KKT factor needed)
  auto hand_off0 = [&](double rho_new, const double (&zc)[NZLs], const double (&yc)[NZLs]) __attribute__((always_inline)) {
    const int b = inst[0];
    int slot = -1;
    if (lane == 0) slot = atomicAdd(io.tail_count, 1);
    slot = __shfl_sync(FULL, slot, 0);
    if (lane == 0) { io.status[b] = ST_HANDOFF; io.iter[b] = it[0]; }
    if (slot < io.tail_capacity) {
      double* ts = io.tail_state + (size_t)slot * (N + 2 * M + 2);
#pragma unroll
      for (int k = 0; k < NXL; ++k) { const int i = lane + 32 * k; if (i < N) ts[i] = x[0][k]; }
#pragma unroll
      for (int k = 0; k < NZL; ++k) { const int j = lane + 32 * k; if (j < M) { ts[N + j] = zc[k]; ts[N + M + j] = yc[k]; } }
      if (lane == 0) { ts[N + 2 * M] = rho_new; ts[N + 2 * M + 1] = (double)it[0]; io.tail_ids[slot] = b; }
    }
  };

  // store_solution / unscale / retrieval for slot 0 (a11, a12)
  auto finish0 = [&](int status, const double (&yc)[NZLs]) __attribute__((always_inline)) {
    const int b = inst[0];
    const bool has_sol = !(status == ST_PINF || status == ST_PINF_INACC || status == ST_DINF ||
                           status == ST_DINF_INACC || status == ST_NONCVX);
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
    for (int k = 0; k < NXL; ++k) {
      const int i = lane + 32 * k;
      if (i < N) {
        const double xv = has_sol ? Dv[i] * x[0][k] : qnan;
        wN[NI * i] = xv;
        if (io.sol_x) io.sol_x[(size_t)b * N + i] = xv;
      }
    }
#pragma unroll
    for (int k = 0; k < NZL; ++k) {
      const int j = lane + 32 * k;
      if (j < M) {
        const double yv = has_sol ? (Ev[j] * yc[k]) * cinv : qnan;
        wN[NI * (N + j)] = yv;
        if (io.sol_y) io.sol_y[(size_t)b * M + j] = yv;
      }
    }
    __syncwarp();
    if (io.prim) {
      const int np = H->n_prim;
      for (int k = lane; k < np; k += LANES) io.prim[(size_t)b * np + k] = wN[NI * U16[H->h_prim + k]];
    }
    if (io.dual) {
      const int nd = H->n_dual;
      for (int k = lane; k < nd; k += LANES) io.dual[(size_t)b * nd + k] = wN[NI * (N + U16[H->h_dual + k])];
    }
    __syncwarp();
    if (lane == 0) {
      double obj = (0.5 * xPx + qx);
      if (st.scaling) obj *= cinv;
      if (status == ST_PINF || status == ST_PINF_INACC) obj = OSQP_INFTY;
      else if (status == ST_DINF || status == ST_DINF_INACC) obj = -OSQP_INFTY;
      else if (status == ST_NONCVX) obj = qnan;
      else obj = (H->is_max ? -1.0 : 1.0) * (obj + H->d_const);
      io.obj_val[b] = obj; io.iter[b] = it[0]; io.status[b] = status;
      io.pri_res[b] = pri_res; io.dua_res[b] = dua_res;
    }
  };

  // fetch instances into slot 0 until one is accepted (or the queue is empty)
  bool need_rhs = true;            // the work vector does not hold the slots' right-hand sides (after a refill / a check iteration)
  auto refill0 = [&]() __attribute__((always_inline)) {
    while (!active[0] && !exhausted) {
      unsigned b = 0;
      if (lane == 0) b = atomicAdd(io.work_counter, 1u);
      b = __shfl_sync(FULL, b, 0);
      if (b >= (unsigned)io.B) { exhausted = true; break; }
      inst[0] = (int)b; it[0] = 0; first[0] = true; need_rhs = true;
      const bool mismatch = load_instance0((int)b);
      // cold start (auxil.c:155-159) or warm start (osqp.c:929-953)
      if (warm) {
#pragma unroll
        for (int k = 0; k < NXL; ++k) {
          const int i = lane + 32 * k;
          if (i < N) { x[0][k] = Dinv[i] * io.x0[(size_t)b * N + i]; wN[NI * i] = x[0][k]; }
        }
        __syncwarp();
        double* wr = io.ws + (size_t)b * (2 * M);        // start values of this instance: z0 = A x0, y0 / rho
#pragma unroll
        for (int k = 0; k < NZL; ++k) {
          const int j = lane + 32 * k;
          if (j < M) {
            wr[j] = ell_dot_col0<NI>(I32 + H->i_ellA + 3 * k, F64, U16, wN, lane);
            wr[M + j] = rinv_of(k) * ((Einv[j] * io.y0[(size_t)b * M + j]) * c);
          }
        }
        __syncwarp();
      } else {
#pragma unroll
        for (int k = 0; k < NXL; ++k) x[0][k] = 0.0;
      }
      if (mismatch || st.max_iter <= 0) {
        double zc[NZLs], yc[NZLs];
        zy0(zc, yc);
        if (mismatch) { hand_off0(rho_in, zc, yc); continue; }     // a constraint changed type: needs its own factor
        double dx0[NXL], dy0[NZLs];                         // degenerate setting: report the start point
#pragma unroll
        for (int k = 0; k < NXL; ++k) dx0[k] = 0.0;
#pragma unroll
        for (int k = 0; k < NZLs; ++k) dy0[k] = 0.0;
        update_info0(zc, yc);
        int status = check_termination0(dx0, dy0, false);
        if (status == ST_UNSOLVED) { status = check_termination0(dx0, dy0, true); if (status == ST_UNSOLVED) status = ST_MAXITER; }
        finish0(status, yc);
        continue;
      }
      active[0] = true;
    }
  };

  // ================================================================ main loop
  // where an element of the KKT right-hand side / solution lives: the warp's own interleaved vectors, or -- tensor-core
  // path -- this warp's pair of columns of the group's w8
  auto wpos = [&](int p) __attribute__((always_inline)) -> double* {
    if constexpr (DMMA) return dc->w8 + w8_off(p, 2 * dc->wg);
    else return wN + NI * p;
  };
  // right-hand side of the next solve from the slots' state (osqp.c:354-358): sigma x - q | z - y / rho
  auto write_rhs = [&]() __attribute__((always_inline)) {
#pragma unroll
    for (int k = 0; k < NXL; ++k) {
      const int i = lane + 32 * k;
      if (i < N) {
        double qv[NI], r[NI];
        tabN(AQ, k, qv);
#pragma unroll
        for (int s = 0; s < NI; ++s) r[s] = sigma * x[s][k] - qv[s];
        MO::st(wpos(PX[32 * k]), r);
      }
    }
#pragma unroll
    for (int k = 0; k < NZL; ++k) {
      const int j = lane + 32 * k;
      if (j < M) {
        double lv[NI], uv[NI], r[NI];
        tabN(AL, k, lv); tabN(AU, k, uv);
#pragma unroll
        for (int s = 0; s < NI; ++s) { double zv, dv; zd_first(s, k, lv[s], uv[s], zv, dv); r[s] = zv - dv; }
        MO::st(wpos(PZ[32 * k]), r);
      }
    }
  };
  // update_x, update_z (+project), update_y (auxil.c:185-225) on the state (x, t), FUSED with the right-hand side of the next
  // solve when RHS (the solution of this one has just been read from the same positions).  FIRST: some slot still holds its
  // start point.
  auto update_all = [&](auto first_tag, auto rhs_tag) __attribute__((always_inline)) {
    constexpr bool FIRST = decltype(first_tag)::value, RHS = decltype(rhs_tag)::value;
#pragma unroll
    for (int k = 0; k < NXL; ++k) {
      const int i = lane + 32 * k;
      if (i < N) {
        double xt[NI], qv[NI], r[NI];
        double* pw = wpos(PX[32 * k]);
        MO::ld(pw, xt);
        if (RHS) tabN(AQ, k, qv);
#pragma unroll
        for (int s = 0; s < NI; ++s) {
          x[s][k] = alpha * xt[s] + (1.0 - alpha) * x[s][k];
          if (RHS) r[s] = sigma * x[s][k] - qv[s];
        }
        if (RHS) MO::st(pw, r);
      }
    }
#pragma unroll
    for (int k = 0; k < NZL; ++k) {
      const int j = lane + 32 * k;
      if (j < M) {
        const double ri = rinv_of(k);
        double nu[NI], lv[NI], uv[NI], r[NI];
        double* pw = wpos(PZ[32 * k]);
        MO::ld(pw, nu);
        tabN(AL, k, lv); tabN(AU, k, uv);
#pragma unroll
        for (int s = 0; s < NI; ++s) {
          double zv, dv;
          if (FIRST) zd_first(s, k, lv[s], uv[s], zv, dv); else zd_of(s, k, lv[s], uv[s], zv, dv);
          const double zt = (zv - dv) + ri * nu[s];
          const double v = alpha * zt + (1.0 - alpha) * zv;
          t[s][k] = v + dv;
          if (RHS) { double zn, dn; zd_of(s, k, lv[s], uv[s], zn, dn); r[s] = zn - dn; }
        }
        if (RHS) MO::st(pw, r);
      }
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < NI; ++s) { it[s] += active[s] ? 1 : 0; first[s] = false; }
  };
  using TrueT = std::integral_constant<bool, true>;
  using FalseT = std::integral_constant<bool, false>;

  int par = 0;
  if constexpr (DMMA) {            // first fill; the group meets before anybody's scratch use can collide with a right-hand side
#pragma unroll 1
    for (int r = 0; r < NI; ++r) { refill0(); rotate_state(); }
    group_bar(dc->bar_id);
  }
  for (;;) {
    bool any_active = false, any_free = false;
#pragma unroll
    for (int s = 0; s < NI; ++s) { any_active |= active[s]; any_free |= !active[s]; }
    if constexpr (!DMMA) {
      // All warps of the CTA run the same ~50 KB of straight-line code per iteration; without this barrier they drift
      // apart and every warp misses the instruction cache on its own (ncu: stall_no_instruction 4.4 cycles/issue).
      // Iteration counts are multiples of check_termination for every instance, so the warps' check iterations coincide.
      if (!__syncthreads_or((int)(!exhausted || any_active))) break;
      if (any_free && !exhausted) {                        // refill empty slots, one rotation at a time
#pragma unroll 1
        for (int r = 0; r < NI; ++r) { refill0(); rotate_state(); }
        any_active = false;
#pragma unroll
        for (int s = 0; s < NI; ++s) any_active |= active[s];
      }
      if (!any_active) continue;    // nothing left for this warp: keep meeting the others at the barrier
    }

    // which slots evaluate their residuals after this iteration
    bool chk[NI], adp[NI];
    bool any_chk = false, any_first = false;
#pragma unroll
    for (int s = 0; s < NI; ++s) {
      const int itn = it[s] + 1;
      const bool can_check = st.check_termination && (itn % st.check_termination == 0);
      adp[s] = active[s] && st.adaptive_rho && st.adaptive_rho_interval && (itn % st.adaptive_rho_interval == 0);
      chk[s] = active[s] && (can_check || itn == st.max_iter);
      any_chk |= chk[s] || adp[s];
      any_first |= first[s];
    }
    if constexpr (DMMA) {           // tell the group: bit 0 = somebody runs the once-per-check routines, bit 1 = somebody is alive
      const int f = (any_chk ? 1 : 0) | (any_active ? 2 : 0);
      if (lane == 0 && f) atomicOr(dc->gflag + par, f);
    }

    // ---- one ADMM iteration for all slots (osqp.c:354-372): rhs (unless the last update left it in place), KKT solve
    if ((DMMA && any_active) || (!DMMA && need_rhs)) write_rhs();
    int gflags = 0;
    if constexpr (DMMA) {
      group_bar(dc->bar_id);                             // right-hand sides and flags of all four warps are in place
      gflags = *reinterpret_cast<volatile int*>(dc->gflag + par);
      if (!(gflags & 2)) break;                          // the whole group is out of work
      dmma_solve(*dc, lane);
      if (dc->wg == 0 && lane == 0) dc->gflag[par] = 0;  // everybody read it before the solve's first barrier
      par ^= 1;
    } else {
      __syncwarp();
#if CPG_FAM_STRAIGHT
      cpg_kkt_solve_gen<NI>(F64, I32, U16, wN, lane);
#endif
    }

    if (!any_chk) {
      if constexpr (DMMA) {
        if (any_active) { if (any_first) update_all(TrueT{}, FalseT{}); else update_all(FalseT{}, FalseT{}); }
        if (gflags & 1) {            // another warp of the group runs its checks in its scratch: frame them with it
          group_bar(dc->bar_id);
          group_bar(dc->bar_id);
        }
      } else {
        if (any_first) update_all(TrueT{}, TrueT{}); else update_all(FalseT{}, TrueT{});
        need_rhs = false;
      }
      continue;
    }

    // ---- same update, keeping the step (dx, dy) for the infeasibility tests of this check iteration
    double dx[NI][NXL], dy[NI][NZLs];
#pragma unroll
    for (int k = 0; k < NXL; ++k) {
      const int i = lane + 32 * k;
#pragma unroll
      for (int s = 0; s < NI; ++s) dx[s][k] = 0.0;
      if (i < N) {
        double xt[NI];
        MO::ld(wpos(PX[32 * k]), xt);
#pragma unroll
        for (int s = 0; s < NI; ++s) {
          const double xn = alpha * xt[s] + (1.0 - alpha) * x[s][k];
          dx[s][k] = xn - x[s][k];
          x[s][k] = xn;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < NZLs; ++k) {
      const int j = lane + 32 * k;
#pragma unroll
      for (int s = 0; s < NI; ++s) dy[s][k] = 0.0;
      if (k < NZL && j < M) {
        const double ri = rinv_of(k), r = rho_of(k);
        double nu[NI], lv[NI], uv[NI];
        MO::ld(wpos(PZ[32 * k]), nu);
        tabN(AL, k, lv); tabN(AU, k, uv);
#pragma unroll
        for (int s = 0; s < NI; ++s) {
          double zv, dv;
          zd_first(s, k, lv[s], uv[s], zv, dv);
          const double zt = (zv - dv) + ri * nu[s];
          const double v = alpha * zt + (1.0 - alpha) * zv;
          t[s][k] = v + dv;
          const double zn = fmin(fmax(t[s][k], lv[s]), uv[s]);
          dy[s][k] = r * (v - zn);                         // delta_y of update_y (auxil.c:215-225)
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < NI; ++s) { it[s] += active[s] ? 1 : 0; first[s] = false; }
    need_rhs = true;
    if constexpr (DMMA) group_bar(dc->bar_id);     // every warp has read its solution: w8 is scratch until the next right-hand side

    // ---- residuals, termination, adaptive-rho decision: slot by slot at position 0
#pragma unroll 1
    for (int r = 0; r < NI; ++r) {
      if (chk[0] || adp[0]) {
        double zc[NZLs], yc[NZLs];
        zy0(zc, yc);
        update_info0(zc, yc);
        int status = ST_UNSOLVED;
        if (chk[0]) status = check_termination0(dx[0], dy[0], false);
        if (status == ST_UNSOLVED && it[0] >= st.max_iter) {          // osqp.c:563-568
          status = check_termination0(dx[0], dy[0], true);
          if (status == ST_UNSOLVED) status = ST_MAXITER;
        }
        if (status != ST_UNSOLVED) { finish0(status, yc); active[0] = false; }
        else if (adp[0]) {                                            // adapt_rho decision (auxil.c:13-74)
          const double pn = s_rp / (fmax(s_z, s_Ax) + DIVISION_TOL);
          const double dn = s_rd / (fmax(fmax(s_q, s_Aty), s_Px) + DIVISION_TOL);
          double rr = rho_in * sqrt(pn / dn);
          rr = fmin(fmax(rr, RHO_MIN), RHO_MAX);
          if (rr > rho_in * st.adaptive_rho_tolerance || rr < rho_in / st.adaptive_rho_tolerance) {
            hand_off0(rr, zc, yc); active[0] = false;
          }
        }
      }
      // rotate everything that is indexed by slot
      rotate_state();
      { const bool c0 = chk[0], a0 = adp[0];
#pragma unroll
        for (int s = 0; s + 1 < NI; ++s) { chk[s] = chk[s + 1]; adp[s] = adp[s + 1]; }
        chk[NI - 1] = c0; adp[NI - 1] = a0; }
#pragma unroll
      for (int k = 0; k < NXL; ++k) { const double v = dx[0][k];
#pragma unroll
        for (int s = 0; s + 1 < NI; ++s) dx[s][k] = dx[s + 1][k];
        dx[NI - 1][k] = v; }
#pragma unroll
      for (int k = 0; k < NZLs; ++k) { const double v = dy[0][k];
#pragma unroll
        for (int s = 0; s + 1 < NI; ++s) dy[s][k] = dy[s + 1][k];
        dy[NI - 1][k] = v; }
    }
    if constexpr (DMMA) {            // refill what terminated (the other kernel does it at the top of its loop), then release w8
#pragma unroll 1
      for (int r = 0; r < NI; ++r) { if (!exhausted) refill0(); rotate_state(); }
      group_bar(dc->bar_id);
    }
  }
}

// registers per thread: the whole register file divided among the CTA's warps, stated as __maxnreg__ (from __launch_bounds__
// alone ptxas settles on 128 for 13 and 14 warps, far below the 152 / 144 that fit)
template <class Fam>
__global__ void __maxnreg__(Fam::MAXREG)
admm_multi_kernel(const uint8_t* __restrict__ blob_g, const BatchIO io, const Settings st) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t total = reinterpret_cast<const CpgBlobHeader*>(blob_g)->total_bytes;
  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (tid == 0) {                      // stage the constants blob with TMA bulk copies
    mbar_expect_tx(&bar, total);
    constexpr uint32_t CHUNK = 32768;
    for (uint32_t off = 0; off < total; off += CHUNK)
      tma_bulk_g2s(smem + off, blob_g + off, (total - off < CHUNK) ? (total - off) : CHUNK, &bar);
  }
  mbar_wait(&bar, 0);
  const CpgBlobHeader* H = reinterpret_cast<const CpgBlobHeader*>(smem);
  const int* I32 = reinterpret_cast<const int*>(smem + H->off_i32);
  const double* F64 = reinterpret_cast<const double*>(smem + H->off_f64);
  const uint16_t* U16 = reinterpret_cast<const uint16_t*>(smem + H->off_u16);
  double* wN = reinterpret_cast<double*>(smem + Fam::BLOB_BYTES_PAD) + (size_t)warp * Fam::MULTI_STRIDE;
  double* bv = wN + Fam::NI * Fam::W_STRIDE;
  solve_multi<Fam, Fam::NI>(H, I32, F64, U16, wN, bv, lane, io, st);
}

#if CPG_FAM_DMMA
// Tensor-core variant: DM_GROUPS groups of four warps per CTA, one persistent CTA per SM.  Shared memory: the compact constants
// blob (no tile schedule) | the DMMA tables (offline/dmma.py, header below) | per group w8 + staging | per warp batched-row
// table | per group two flag words.  Both blobs are staged by TMA bulk copies.
struct CpgDmmaHeader { int total_bytes, n_tiles, off_hdr, off_rl, off_items, off_vals, off_jobs, max_rounds; };

template <class Fam>
__global__ void __launch_bounds__(Fam::DM_GROUPS * 128, 1)
admm_dmma_kernel(const uint8_t* __restrict__ blob_g, const uint8_t* __restrict__ dblob_g, const BatchIO io, const Settings st) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, grp = warp >> 2;
  const uint32_t total = reinterpret_cast<const CpgBlobHeader*>(blob_g)->total_bytes;
  const uint32_t dtotal = reinterpret_cast<const CpgDmmaHeader*>(dblob_g)->total_bytes;
  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, total + dtotal);
    constexpr uint32_t CHUNK = 32768;
    for (uint32_t off = 0; off < total; off += CHUNK)
      tma_bulk_g2s(smem + off, blob_g + off, (total - off < CHUNK) ? (total - off) : CHUNK, &bar);
    for (uint32_t off = 0; off < dtotal; off += CHUNK)
      tma_bulk_g2s(smem + Fam::CBLOB_BYTES_PAD + off, dblob_g + off, (dtotal - off < CHUNK) ? (dtotal - off) : CHUNK, &bar);
  }
  int* gflags = reinterpret_cast<int*>(smem + Fam::CBLOB_BYTES_PAD + Fam::DBLOB_BYTES_PAD +
                                       (size_t)Fam::DM_GROUPS * (Fam::DM_W8 + Fam::DM_STAGE) * 8 + (size_t)Fam::DM_GROUPS * 4 * Fam::DM_BV * 8);
  if (tid < 2 * Fam::DM_GROUPS) gflags[tid] = 0;
  __syncthreads();
  mbar_wait(&bar, 0);
  const CpgBlobHeader* H = reinterpret_cast<const CpgBlobHeader*>(smem);
  const int* I32 = reinterpret_cast<const int*>(smem + H->off_i32);
  const double* F64 = reinterpret_cast<const double*>(smem + H->off_f64);
  const uint16_t* U16 = reinterpret_cast<const uint16_t*>(smem + H->off_u16);
  const uint8_t* db = smem + Fam::CBLOB_BYTES_PAD;
  const CpgDmmaHeader* DH = reinterpret_cast<const CpgDmmaHeader*>(db);
  double* gbase = reinterpret_cast<double*>(smem + Fam::CBLOB_BYTES_PAD + Fam::DBLOB_BYTES_PAD) + (size_t)grp * (Fam::DM_W8 + Fam::DM_STAGE);
  DmmaCtx dc;
  dc.w8 = gbase; dc.stage = gbase + Fam::DM_W8;
  dc.items = reinterpret_cast<const uint4*>(db + DH->off_items);
  dc.vals = reinterpret_cast<const double*>(db + DH->off_vals);
  dc.tile_hdr = reinterpret_cast<const int4*>(db + DH->off_hdr);
  dc.round_len = reinterpret_cast<const uint16_t*>(db + DH->off_rl);
  dc.jobs = reinterpret_cast<const uint32_t*>(db + DH->off_jobs);
  dc.gflag = gflags + 2 * grp;
  dc.n_tiles = DH->n_tiles; dc.wg = warp & 3; dc.bar_id = 1 + grp;
  double* bv = reinterpret_cast<double*>(smem + Fam::CBLOB_BYTES_PAD + Fam::DBLOB_BYTES_PAD) +
               (size_t)Fam::DM_GROUPS * (Fam::DM_W8 + Fam::DM_STAGE) + (size_t)warp * Fam::DM_BV;
  double* scratch = dc.w8 + (size_t)dc.wg * (Fam::DM_W8 / 4);      // >= 2 * W_STRIDE doubles
  solve_multi<Fam, 2, true>(H, I32, F64, U16, scratch, bv, lane, io, st, &dc);
}
#endif

}  // namespace cpgb200
