// ipm_kernel.cuh -- batched Mehrotra predictor-corrector interior-point method for one SOCP family (IPM-CUDA backend).
//
// One CTA solves one problem instance at a time (persistent CTAs pull instances from a global counter): the iterate,
// the NT scalings, the numeric LDL' factor of the stretched KKT matrix and all work vectors live in shared memory, the
// constant index tables of the family are staged once per CTA.  The algorithm is the one the reference runs per
// instance on the host -- ECOS 2.0.8 as driven by cvxpygen's generated code (cvxpygen/solvers/ecos.py:88-117) --
// re-organised into barrier-separated phases:
//
//   init                      ecos/src/ecos.c:260-452          kkt_init        ecos/src/kkt.c:373-445
//   computeResiduals          ecos/src/ecos.c:455-499          updateStatistics ecos/src/ecos.c:502-545
//   checkExitConditions       ecos/src/ecos.c:179-257          compareStatistics / best iterate :61-149
//   updateScalings / scale    ecos/src/cone.c:138-234,276-305  kkt_update      ecos/src/kkt.c:271-357
//   kkt_factor (LDL_numeric2 + dynamic regularisation)         ecos/external/ldl/src/ldl.c:266-360
//   kkt_solve (iterative refinement)  ecos/src/kkt.c:87-265    scale2add       ecos/src/cone.c:313-400
//   RHS_affine / RHS_combined ecos/src/ecos.c:648-757          lineSearch      ecos/src/ecos.c:947-1046
//   conicProduct / conicDivision ecos/src/cone.c:452-513       main loop / backscale ecos/src/ecos.c:1075-1607,1051-1070
//
// Every sparse phase (one level of the numeric LDL', one level of a triangular solve, a product with the constant part of
// K) runs as a GATHER PLAN (cvxpygen_b200/offline/gather.py): each target is owned by one lane group, entries are read
// lane-interleaved from tables, long rows are summed with butterfly shuffles -- no atomics, fixed summation order.
//
// Index space: k = [x (N) | y (P) | z stretched (MT = M + 2 NSOC)]; every second-order cone of size d occupies d + 2
// consecutive rows, the last two carrying the sparse representation of its scaling (ecos/src/preproc.c:77-330).
// Tables come from cvxpygen_b200/offline/socp_setup.py; sizes and offsets are compile-time constants (cpg_ipm_family.h).
//
// The same source compiles as plain host C++ with -DCPG_IPM_HOST_EMU: every phase then runs its threads one after the
// other.  That build exists for the CPU tests of the phase logic only (tests/emu); it is not reachable from the product.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef CPG_IPM_HOST_EMU
#include <algorithm>
#include <cstring>
#include <vector>
#define IPM_FN inline
#define IPM_NOINLINE inline
#define IPM_HD
#define IPM_CONST static const
namespace cpgipm { static long g_count[8] = {0, 0, 0, 0, 0, 0, 0, 0}; }     // emulation statistics: solves, factors, barriers
#define IPM_COUNT(i) (++cpgipm::g_count[i])
#else
#define IPM_FN __device__ __forceinline__
#define IPM_NOINLINE __device__ __noinline__
#define IPM_HD __host__ __device__
#define IPM_CONST __constant__ const
#define IPM_COUNT(i) ((void)0)
#endif

#ifndef IPM_MATPAR
#define IPM_MATPAR 0
#define IPM_NEMAP 0
#endif

namespace cpgipm {

// ---- constants of the reference (ecos/include/ecos.h:45-75)
constexpr double kGamma = 0.99, kDeltaStat = 7e-8, kDelta = 2e-7, kEps = 1e-13;
constexpr int kNitref = 9;
constexpr double kIrErrFact = 6.0, kLinsysAcc = 1e-14;
constexpr double kSigmaMin = 1e-4, kSigmaMax = 1.0, kStepMin = 1e-6, kStepMax = 0.999, kSafeguard = 500.0;
enum { kOptimal = 0, kPinf = 1, kDinf = 2, kInaccOffset = 10, kMaxit = -1, kNumerics = -2, kOutcone = -3, kFatal = -7,
       kNotConverged = -87 };

struct IpmSettings {
  int maxit;                                  // 100   (cvxpygen/solvers/ecos.py:60-68)
  int pad_;
  double feastol, abstol, reltol;             // 1e-8
  double feastol_inacc, abstol_inacc, reltol_inacc;   // 1e-4, 5e-5, 5e-5
};

struct IpmIO {
  int B;
  const double* params;      // (B, NPB)
  double* prim;              // (B, NPRIM)
  double* dual;              // (B, NDUAL)
  double* sol_x;             // (B, N) or null
  double* sol_y;             // (B, P) or null
  double* sol_z;             // (B, M) or null
  double* sol_s;             // (B, M) or null
  double* obj_val; int* iter; int* status; double* pri_res; double* dua_res;
  double* best;              // per-CTA scratch: gridDim.x * BEST_STRIDE doubles -- the best iterate (NK + MT) and, with per-instance
                             // matrices, the instance's output scalings 1/xe, 1/Ae, 1/Ge in k-space (NK)
  int* counter;              // work counter
};

constexpr int T = CPG_IPM_THREADS;
constexpr int NWARP = T / 32;
constexpr int N = IPM_N, P = IPM_P, M = IPM_M, L = IPM_L, NSOC = IPM_NSOC, MT = IPM_MT, NK = IPM_NK, ZOFF = IPM_ZOFF;
constexpr int NW = IPM_NW, DG0 = IPM_DG0, TT0 = IPM_TT0, NS = IPM_NS, NT = IPM_NT, NLW = IPM_NLW;
constexpr int NNZM = IPM_NNZM, NPB = IPM_NPB, NMAP = IPM_NMAP, NPRIM = IPM_NPRIM, NDUAL = IPM_NDUAL, QTOT = IPM_QTOT;
constexpr int CONE_D = L + NSOC;              // degree of the cone (w->D)
constexpr int PT = (NK + T - 1) / T;          // elements of a k-space vector owned by one thread
constexpr int BEST_STRIDE = NK + MT + (IPM_MATPAR ? NK : 0);      // doubles of IpmIO::best per CTA
constexpr int NCR = MT - L;                   // stretched rows of the second-order cones
static_assert(IPM_THREADS == CPG_IPM_THREADS, "the gather plans were dealt for another CTA width: regenerate the family");
static_assert(NT <= 32, "the dense tail block is handled by one warp");

// ---- shared-memory layout (in doubles, then the u32 / u16 tables).  S carries one extra slot that stays zero (the null
// entry of the plans points at it); the inverse pivots live in the diagonal slots of S (dinv(k) = S[DG0 + k]).
constexpr int O_XYZ = 0, O_CBH = O_XYZ + NK, O_SV = O_CBH + NK, O_LAM = O_SV + MT, O_V = O_LAM + MT, O_W = O_V + L,
              O_Q = O_W + L, O_SC = O_Q + QTOT, O_S = O_SC + 8 * (NSOC > 0 ? NSOC : 1),
              O_RHS = O_S + NS + 1, O_PX = O_RHS + NK, O_E = O_PX + NK, O_SOL1 = O_E + NK, O_RZ = O_SOL1 + NK,
              O_DSW = O_RZ + MT, O_RED = O_DSW + MT, O_CONE = O_RED + 2 * NWARP * 16,
              O_CE = O_CONE + 4 * (NSOC > 0 ? NSOC : 1), O_TW = O_CE + (NCR > 0 ? NCR : 1),
              O_AG = (O_TW + 2 * 32 + 1) & ~1, O_F64_END = O_AG + NNZM + 1;      // O_AG even: 16-byte aligned target of the bulk copy
constexpr int U32_COUNT = (IPM_SB_U16_OFF - IPM_SB_U32_OFF) / 4;
constexpr int U16_COUNT = (IPM_SB_BYTES - IPM_SB_U16_OFF) / 2;
constexpr size_t SMEM_BYTES = size_t(O_F64_END) * 8 + size_t(U32_COUNT) * 4 + size_t(U16_COUNT) * 2 + 16;
enum { SC_ETA2 = 0, SC_ETA, SC_A, SC_D1, SC_U0, SC_U1, SC_V1, SC_W };

// Shared-memory view.  On the device every array sits at a compile-time offset of the dynamic shared-memory window, so an
// access costs no pointer register; the host emulation carries the base of a heap buffer instead.
#ifndef CPG_IPM_HOST_EMU
extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
struct Sm {
#ifdef CPG_IPM_HOST_EMU
  unsigned char* base_;
  IPM_FN unsigned char* base() const { return base_; }
#else
  IPM_FN unsigned char* base() const { return smem_raw; }
#endif
  IPM_FN double* f64(int off) const { return reinterpret_cast<double*>(base()) + off; }
  IPM_FN const uint32_t* u32(int off) const { return reinterpret_cast<const uint32_t*>(f64(O_F64_END)) + off; }
  IPM_FN const uint16_t* u16(int off) const { return reinterpret_cast<const uint16_t*>(u32(U32_COUNT)) + off; }
  IPM_FN double* xyz() const { return f64(O_XYZ); }   IPM_FN double* cbh() const { return f64(O_CBH); }
  IPM_FN double* sv() const { return f64(O_SV); }     IPM_FN double* lam() const { return f64(O_LAM); }
  IPM_FN double* v() const { return f64(O_V); }       IPM_FN double* w() const { return f64(O_W); }
  IPM_FN double* q() const { return f64(O_Q); }       IPM_FN double* sc() const { return f64(O_SC); }
  IPM_FN double* S() const { return f64(O_S); }       IPM_FN double* rhs() const { return f64(O_RHS); }
  IPM_FN double* px() const { return f64(O_PX); }     IPM_FN double* e() const { return f64(O_E); }
  IPM_FN double* sol1() const { return f64(O_SOL1); } IPM_FN double* rz() const { return f64(O_RZ); }
  IPM_FN double* dsw() const { return f64(O_DSW); }   IPM_FN double* red() const { return f64(O_RED); }
  IPM_FN double* cone() const { return f64(O_CONE); } IPM_FN double* ce() const { return f64(O_CE); }
  IPM_FN double* tw() const { return f64(O_TW); }     IPM_FN double* ag() const { return f64(O_AG); }
  IPM_FN const uint32_t* mv_e() const { return u32(IPM_E_MV); }
  IPM_FN const uint32_t* fw_e() const { return u32(IPM_E_FW); }
  IPM_FN const uint32_t* bw_e() const { return u32(IPM_E_BW); }
  IPM_FN const uint16_t* mv_d() const { return u16(IPM_H_MV_D); }
  IPM_FN const uint16_t* fw_d() const { return u16(IPM_H_FW_D); }
  IPM_FN const uint16_t* bw_d() const { return u16(IPM_H_BW_D); }
  IPM_FN const uint16_t* socv() const { return u16(IPM_H_SOCV); }
  IPM_FN const uint16_t* socu() const { return u16(IPM_H_SOCU); }
  IPM_FN const uint16_t* tail_k() const { return u16(IPM_H_TAIL_K); }
  IPM_FN const uint16_t* perm() const { return u16(IPM_H_PERM); }
  IPM_FN const uint16_t* l0mask() const { return u16(IPM_H_L0MASK); }
  IPM_FN int* flag() const { return reinterpret_cast<int*>(const_cast<uint16_t*>(u16(U16_COUNT + (U16_COUNT & 1)))); }
};

struct Gm {                 // global constant tables
  const double *Sbase, *cbh_base, *unscale, *map_v;
  const int *map_t, *map_p, *prim_idx, *dual_idx;
#if IPM_MATPAR      // per-instance G / A values: raw entry = ent_base + emap_v * theta[emap_p]; entry e couples k-rows mr_t[e], mr_s[e] (slot ag_slot[e])
  const double *ent_base, *emap_v;
  const int *emap_t, *emap_p, *mr_t, *mr_s, *ag_slot;
#endif
  const unsigned long long* op_e;      // factorisation plan: entries (slot a | slot b << 16 | diagonal slot << 32)
  const uint32_t* op_d;                // and descriptors
};

IPM_FN Gm make_gm(const unsigned char* g) {
  Gm r;
  const double* f = reinterpret_cast<const double*>(g);
  r.Sbase = f + IPM_G_SBASE; r.cbh_base = f + IPM_G_CBH_BASE; r.unscale = f + IPM_G_UNSCALE; r.map_v = f + IPM_G_MAP_V;
  const int* i = reinterpret_cast<const int*>(g + IPM_GB_I32_OFF);
  r.map_t = i + IPM_GI_MAP_T; r.map_p = i + IPM_GI_MAP_P; r.prim_idx = i + IPM_GI_PRIM_IDX; r.dual_idx = i + IPM_GI_DUAL_IDX;
#if IPM_MATPAR
  r.ent_base = f + IPM_G_ENT_BASE; r.emap_v = f + IPM_G_EMAP_V;
  r.emap_t = i + IPM_GI_EMAP_T; r.emap_p = i + IPM_GI_EMAP_P; r.mr_t = i + IPM_GI_MR_T; r.mr_s = i + IPM_GI_MR_S; r.ag_slot = i + IPM_GI_AG_SLOT;
#endif
  r.op_e = reinterpret_cast<const unsigned long long*>(g + IPM_GB_OPS_OFF);
  r.op_d = reinterpret_cast<const uint32_t*>(g + IPM_GB_OPD_OFF);
  return r;
}

IPM_FN double safediv(double x, double y) { return y < kEps ? x / kEps : x / y; }

// ----------------------------------------------------------------------------------------------------------------------
// execution model: phases, reductions, per-cone warps
#ifdef CPG_IPM_HOST_EMU
IPM_FN void atomic_add(double* p, double v) { *p += v; }
IPM_FN void atomic_max_nonneg(double* p, double v) { if (v > *p) *p = v; }
template <class F> IPM_FN void phase(F&& f) { IPM_COUNT(2); for (int t = 0; t < T; ++t) f(t); }
IPM_FN void sync_phase() { IPM_COUNT(2); }
// KS sums and KM maxima over all threads; f(tid, s, m) accumulates with += and fmax
template <int KS, int KM, class F> IPM_FN void phase_red(Sm&, int&, double* s, double* m, F&& f) {
  for (int k = 0; k < KS; ++k) s[k] = 0.0;
  for (int k = 0; k < KM; ++k) m[k] = -INFINITY;
  for (int t = 0; t < T; ++t) f(t, s, m);
}
struct WarpOps {
  template <class F> double sum(int lo, int hi, F&& f) const { double s = 0; for (int i = lo; i < hi; ++i) s += f(i); return s; }
  template <class F> void each(int lo, int hi, F&& f) const { for (int i = lo; i < hi; ++i) f(i); }
  bool leader() const { return true; }
  void sync() const {}
};
// thread part f(tid) and cone part g(cone, WarpOps) of one phase (independent of each other)
template <class F, class G> IPM_FN void phase_cones(F&& f, G&& g) {
  for (int t = 0; t < T; ++t) f(t);
  for (int c = 0; c < NSOC; ++c) g(c, WarpOps());
}
template <int KS, int KM, class F, class G>
IPM_FN void phase_red_cones(Sm& sm, int& rb, double* s, double* m, F&& f, G&& g) {
  phase_red<KS, KM>(sm, rb, s, m, f);
  for (int c = 0; c < NSOC; ++c) g(c, WarpOps());
}
template <class G> IPM_FN void cones_then_sync(G&& g) { for (int c = 0; c < NSOC; ++c) g(c, WarpOps()); }
template <int CNT> struct PerThread {
  std::vector<double> a; PerThread() : a(size_t(T) * CNT, 0.0) {}
  double& at(int tid, int j) { return a[size_t(tid) * CNT + j]; }
};
#else
IPM_FN void atomic_add(double* p, double v) { atomicAdd(p, v); }
// max of NON-NEGATIVE doubles: their bit patterns order like unsigned integers (exact, order-independent)
IPM_FN void atomic_max_nonneg(double* p, double v) {
#ifdef CPG_SIMT_HOST_EMU
  if (v > *p) *p = v;
#else
  atomicMax(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(__double_as_longlong(v)));
#endif
}
template <class F> IPM_FN void phase(F&& f) { f(int(threadIdx.x)); __syncthreads(); }
IPM_FN void sync_phase() { __syncthreads(); }
IPM_FN double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Maximum over the warp with fmax semantics (NaN entries are ignored), on an order-preserving integer image of the
// doubles: two REDUX.MAX (high word signed, then low word among the lanes that hold the largest high word) instead of
// five shuffle + DSETP/SEL rounds -- double has no single-instruction max on this architecture.
IPM_FN double warp_max(double v) {
  if (!(v == v)) v = -INFINITY;
  long long key = __double_as_longlong(v);
  key ^= (key >> 63) & 0x7fffffffffffffffLL;
  const int hi = int(key >> 32);
  const unsigned lo = unsigned(key);
  const int mhi = __reduce_max_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
  key = (static_cast<long long>(mhi) << 32) | static_cast<long long>(mlo);
  key ^= (key >> 63) & 0x7fffffffffffffffLL;
  return __longlong_as_double(key);
}
template <int KS, int KM, class F> IPM_FN void phase_red(Sm& sm, int& rb, double* s, double* m, F&& f) {
  static_assert(KS + KM <= 16 && NWARP <= 32, "reduction scratch holds 16 values per warp");
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int k = 0; k < KS; ++k) s[k] = 0.0;
#pragma unroll
  for (int k = 0; k < KM; ++k) m[k] = -INFINITY;
  f(tid, s, m);
  double* buf = sm.red() + rb * (NWARP * 16);
  rb ^= 1;
#pragma unroll
  for (int k = 0; k < KS; ++k) { double v = warp_sum(s[k]); if (lane == 0) buf[wid * 16 + k] = v; }
#pragma unroll
  for (int k = 0; k < KM; ++k) { double v = warp_max(m[k]); if (lane == 0) buf[wid * 16 + KS + k] = v; }
  __syncthreads();
#pragma unroll
  if constexpr (KS > 8 && NWARP % 2 == 0) {
    // many sums (the statistics of an iteration): lane (k, half) adds the partials of half of the warps, one shuffle joins
    // the halves, KS broadcasts hand every thread every total -- instead of KS * NWARP dependent loads and adds per thread
    const int k = lane & 15, half = lane >> 4;
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < NWARP / 2; ++w) v += buf[(half * (NWARP / 2) + w) * 16 + k];
    v += __shfl_xor_sync(0xffffffffu, v, 16);
#pragma unroll
    for (int j = 0; j < KS; ++j) s[j] = __shfl_sync(0xffffffffu, v, j);
  } else {
#pragma unroll
    for (int k = 0; k < KS; ++k) { double v = 0.0; for (int w = 0; w < NWARP; ++w) v += buf[w * 16 + k]; s[k] = v; }
  }
#pragma unroll
  for (int k = 0; k < KM; ++k) m[k] = warp_max(lane < NWARP ? buf[lane * 16 + KS + k] : -INFINITY);
}
struct WarpOps {
  int lane;
  template <class F> IPM_FN double sum(int lo, int hi, F&& f) const {
    double s = 0; for (int i = lo + lane; i < hi; i += 32) s += f(i); return warp_sum(s);
  }
  template <class F> IPM_FN void each(int lo, int hi, F&& f) const { for (int i = lo + lane; i < hi; i += 32) f(i); }
  IPM_FN bool leader() const { return lane == 0; }
  IPM_FN void sync() const { __syncwarp(); }
};
template <class G> IPM_FN void run_cones(G&& g) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // cones go to the LAST warps first so that they overlap with the (front-loaded) elementwise work of the others
  for (int c = NWARP - 1 - wid; c < NSOC; c += NWARP) g(c, WarpOps{lane});
}
template <class F, class G> IPM_FN void phase_cones(F&& f, G&& g) {
  f(int(threadIdx.x)); run_cones(g); __syncthreads();
}
template <class G> IPM_FN void cones_then_sync(G&& g) { run_cones(g); __syncthreads(); }
template <int KS, int KM, class F, class G>
IPM_FN void phase_red_cones(Sm& sm, int& rb, double* s, double* m, F&& f, G&& g) {
  run_cones(g);
  phase_red<KS, KM>(sm, rb, s, m, f);
}
template <int CNT> struct PerThread {
  double a[CNT];
  IPM_FN double& at(int, int j) { return a[j]; }
};
#endif

template <class F> IPM_FN void each_k(int tid, int n, F&& f) { for (int i = tid; i < n; i += T) f(i); }
// elementwise work of a phase whose second-order cones are handled by the last warps (run_cones): those warps are left
// out here, so that the cone arithmetic overlaps the elementwise part instead of following it
constexpr int CONE_WARPS = NSOC < NWARP - 1 ? NSOC : NWARP - 1;
#ifdef CPG_IPM_HOST_EMU
template <class F> IPM_FN void each_k_nc(int tid, int n, F&& f) { each_k(tid, n, f); }
#else
template <class F> IPM_FN void each_k_nc(int tid, int n, F&& f) {
  constexpr int TN = T - 32 * CONE_WARPS;
  if (tid < TN) for (int i = tid; i < n; i += TN) f(i);
}
#endif

IPM_CONST int kSocSo[NSOC > 0 ? NSOC : 1] = IPM_SOC_SO;      // stretched z offset of each cone
IPM_CONST int kSocD[NSOC > 0 ? NSOC : 1] = IPM_SOC_D;        // cone sizes
IPM_CONST int kSocQo[NSOC > 0 ? NSOC : 1] = IPM_SOC_QO;      // offset of q in sm.q()
IPM_CONST int kSocVo[NSOC > 0 ? NSOC : 1] = IPM_SOC_VO;      // offset in socv
IPM_CONST int kSocUo[NSOC > 0 ? NSOC : 1] = IPM_SOC_UO;      // offset in socu
IPM_CONST int kLevLo[NLW + 1] = IPM_LEV_LO;

// ---- gather plans.  Run-time tables: per (round, warp) one word  entry offset | trip count << 20.  Compile-time
// description (family header): the rounds of each phase and, per round, the longest cell and the deepest butterfly of any
// warp -- every loop of a round is unrolled to those bounds, so a round is straight-line code.
IPM_CONST unsigned kMvWr[] = IPM_MV_WR;
IPM_CONST unsigned kFwWr[] = IPM_FW_WR;
IPM_CONST unsigned kBwWr[] = IPM_BW_WR;
IPM_CONST unsigned kOpWr[] = IPM_OP_WR;
struct MvShape { static constexpr int ph[] = IPM_MV_PH; static constexpr int lmax[] = IPM_MV_LMAX; static constexpr int lmin[] = IPM_MV_LMIN; static constexpr int smax[] = IPM_MV_SMAX; };
struct FwShape { static constexpr int ph[] = IPM_FW_PH; static constexpr int lmax[] = IPM_FW_LMAX; static constexpr int lmin[] = IPM_FW_LMIN; static constexpr int smax[] = IPM_FW_SMAX; };
struct BwShape { static constexpr int ph[] = IPM_BW_PH; static constexpr int lmax[] = IPM_BW_LMAX; static constexpr int lmin[] = IPM_BW_LMIN; static constexpr int smax[] = IPM_BW_SMAX; };
struct OpShape { static constexpr int ph[] = IPM_OP_PH; static constexpr int lmax[] = IPM_OP_LMAX; static constexpr int lmin[] = IPM_OP_LMIN; static constexpr int smax[] = IPM_OP_SMAX; };

// entries and descriptors carry BYTE offsets (index * 8) into the f64 arrays: one add-free LDS per operand
IPM_FN double at(const double* base, unsigned byte_off) { return *reinterpret_cast<const double*>(reinterpret_cast<const char*>(base) + byte_off); }
template <int I> struct IC { static constexpr int value = I; };
template <int I, int E_, class F> IPM_FN void static_for(F&& f) {
  if constexpr (I < E_) { f(IC<I>{}); static_for<I + 1, E_>(f); }
}

// One plan = entry table (u32: two u16 fields, or u64: up to four), descriptor table (u16 with an 11-bit target, or u32
// with a 16-bit target) and the words above.  run<PH>(value, commit): for every cell of phase PH
//   acc = sum_j value(entry_j);  butterfly over the cell group;  the group leader calls commit(target, flag, acc).
template <class Shape, class E, class D, int TB> struct Plan {
  const E* ent; const D* desc; const unsigned* wr;
  template <int R, class V, class C> IPM_FN void round(V&& value, C&& commit) const {
    constexpr int LMAX = Shape::lmax[R], LMIN = Shape::lmin[R], SMAX = Shape::smax[R];      // LMIN: shortest warp of the round
#ifdef CPG_IPM_HOST_EMU
    for (int w = 0; w < NWARP; ++w) {
      const unsigned m = wr[R * NWARP + w];
      const int base = int(m & 0xfffffu), len = int(m >> 20);
      double acc[32]; unsigned d[32]; int g[32];
      for (int l = 0; l < 32; ++l) {
        d[l] = desc[R * T + w * 32 + l]; g[l] = 1 << ((d[l] >> TB) & 7u); acc[l] = 0.0;
        for (int j = 0; j < LMAX; ++j) if (j < len) acc[l] += value(ent[base + j * 32 + l]);
      }
      for (int s = 0; s < SMAX; ++s) {
        const int o = 1 << s;
        double nx[32];
        for (int l = 0; l < 32; ++l) nx[l] = o < g[l] ? acc[l] + acc[l ^ o] : acc[l];
        for (int l = 0; l < 32; ++l) acc[l] = nx[l];
      }
      for (int l = 0; l < 32; ++l)
        if (((d[l] >> (TB + 3)) & 1u) && (l & (g[l] - 1)) == 0) commit(int(d[l] & ((1u << TB) - 1u)), int((d[l] >> (TB + 4)) & 1u), acc[l]);
    }
#else
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned m = wr[R * NWARP + wid];
    const unsigned d = desc[R * T + tid];
    const E* e = ent + (m & 0xfffffu) + lane;
    const int len = int(m >> 20);                  // warp-uniform
    double acc = 0.0;
#pragma unroll
    for (int j0 = 0; j0 < LMAX; j0 += 8) {         // table reads of a chunk first: they do not depend on the arithmetic
      E ev[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) if (j0 + j < LMAX && (j0 + j < LMIN || j0 + j < len)) ev[j] = e[(j0 + j) * 32];
#pragma unroll
      for (int j = 0; j < 8; ++j) if (j0 + j < LMAX && (j0 + j < LMIN || j0 + j < len)) acc += value(ev[j]);
    }
    const int g = 1 << ((d >> TB) & 7u);
#pragma unroll
    for (int s = 0; s < SMAX; ++s) {
      const double v = __shfl_xor_sync(0xffffffffu, acc, 1 << s);
      if ((1 << s) < g) acc += v;
    }
    if (((d >> (TB + 3)) & 1u) && (lane & (g - 1)) == 0) commit(int(d & ((1u << TB) - 1u)), int((d >> (TB + 4)) & 1u), acc);
#endif
  }
  // total unrolled trip count of a phase (compile time)
  IPM_HD static constexpr int phase_entries(int ph_) {
    int n = 0;
    for (int r = Shape::ph[ph_]; r < Shape::ph[ph_ + 1]; ++r) n += Shape::lmax[r];
    return n;
  }
  IPM_HD static constexpr int phase_lmax(int ph_) {
    int n = 1;
    for (int r = Shape::ph[ph_]; r < Shape::ph[ph_ + 1]; ++r) n = Shape::lmax[r] > n ? Shape::lmax[r] : n;
    return n;
  }
#ifndef CPG_IPM_HOST_EMU
  // The rounds of one phase are independent (no target of a phase is a source in it), so a short multi-round phase runs
  // stage by stage ACROSS its rounds -- all table reads, then all operand reads and sums, then the butterflies, then the
  // commits: one warp has several dependent chains in flight instead of one (the kernel is latency bound).
  template <int PH, class V, class C> IPM_FN void run_staged(V&& value, C&& commit) const {
    constexpr int R0 = Shape::ph[PH], NR = Shape::ph[PH + 1] - R0, LM = phase_lmax(PH);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned m[NR], d[NR];
    E ev[NR][LM];
    double acc[NR];
    static_for<0, NR>([&](auto r) {
      constexpr int i = decltype(r)::value;
      m[i] = wr[(R0 + i) * NWARP + wid];
      d[i] = desc[(R0 + i) * T + tid];
    });
    static_for<0, NR>([&](auto r) {
      constexpr int i = decltype(r)::value, LMAX = Shape::lmax[R0 + i], LMIN = Shape::lmin[R0 + i];
      const E* e = ent + (m[i] & 0xfffffu) + lane;
      const int len = int(m[i] >> 20);
#pragma unroll
      for (int j = 0; j < LMAX; ++j) if (j < LMIN || j < len) ev[i][j] = e[j * 32];
    });
    static_for<0, NR>([&](auto r) {
      constexpr int i = decltype(r)::value, LMAX = Shape::lmax[R0 + i], LMIN = Shape::lmin[R0 + i];
      const int len = int(m[i] >> 20);
      double a = 0.0;
#pragma unroll
      for (int j = 0; j < LMAX; ++j) if (j < LMIN || j < len) a += value(ev[i][j]);
      acc[i] = a;
    });
    static_for<0, NR>([&](auto r) {
      constexpr int i = decltype(r)::value, SMAX = Shape::smax[R0 + i];
      const int g = 1 << ((d[i] >> TB) & 7u);
#pragma unroll
      for (int s = 0; s < SMAX; ++s) {
        const double v = __shfl_xor_sync(0xffffffffu, acc[i], 1 << s);
        if ((1 << s) < g) acc[i] += v;
      }
    });
    static_for<0, NR>([&](auto r) {
      constexpr int i = decltype(r)::value;
      const int g = 1 << ((d[i] >> TB) & 7u);
      if (((d[i] >> (TB + 3)) & 1u) && (lane & (g - 1)) == 0)
        commit(int(d[i] & ((1u << TB) - 1u)), int((d[i] >> (TB + 4)) & 1u), acc[i]);
    });
  }
#endif
  template <int PH, class V, class C> IPM_FN void run(V&& value, C&& commit) const {
#ifndef CPG_IPM_HOST_EMU
    constexpr int NR = Shape::ph[PH + 1] - Shape::ph[PH];
    if constexpr (NR >= 2 && phase_entries(PH) * int(sizeof(E)) <= 96) { run_staged<PH>(value, commit); return; }
#endif
    static_for<Shape::ph[PH], Shape::ph[PH + 1]>([&](auto r) { this->template round<decltype(r)::value>(value, commit); });
  }
};
using PlanMv = Plan<MvShape, uint32_t, uint16_t, 11>;             // tables in shared memory
using PlanFw = Plan<FwShape, uint32_t, uint16_t, 11>;
using PlanBw = Plan<BwShape, uint32_t, uint16_t, 11>;
using PlanOp = Plan<OpShape, unsigned long long, uint32_t, 16>;   // tables in global memory (read once per factorisation)

// ----------------------------------------------------------------------------------------------------------------------
struct Solver {
  Sm sm; Gm gm; IpmSettings stg;
  int rb;                                   // reduction scratch toggle

  IPM_FN PlanMv plan_mv() const { return PlanMv{sm.mv_e(), sm.mv_d(), kMvWr}; }
  IPM_FN PlanFw plan_fw() const { return PlanFw{sm.fw_e(), sm.fw_d(), kFwWr}; }
  IPM_FN PlanBw plan_bw() const { return PlanBw{sm.bw_e(), sm.bw_d(), kBwWr}; }
  IPM_FN PlanOp plan_op() const { return PlanOp{gm.op_e, gm.op_d, kOpWr}; }

  IPM_FN double sign_of(int k) const {      // Sign vector of createKKT_U (preproc.c:135-170)
    if (k < N) return 1.0;
    for (int c = 0; c < NSOC; ++c) if (k == ZOFF + kSocSo[c] + kSocD[c] + 1) return 1.0;
    return -1.0;
  }
  // pivot of column k with the dynamic regularisation of LDL_numeric2 (ldl.c:336-350); its inverse replaces it in S
  IPM_FN double inv_pivot(int k, double d) const {
    const double sg = sign_of(k);
    if (sg * d <= kEps) d = sg * kDelta;
    return 1.0 / d;
  }

  // ---- numeric factorisation: S holds the KKT values on entry; on exit the column-scaled factor (S_ij = L_ij D_j), the
  // inverse pivots in the diagonal slots and the inverse of the tail's unit triangle in the tail block
  IPM_FN void tail_factor(int tid);
  IPM_NOINLINE void factor() {
    IPM_COUNT(1);
    phase([&](int tid) {
      for (int p = kLevLo[0] + tid; p < kLevLo[1]; p += T) { const int k = sm.perm()[p]; sm.S()[DG0 + k] = inv_pivot(k, sm.S()[DG0 + k]); }
    });
    const PlanOp po = plan_op();
    static_for<0, NLW>([&](auto lv) {
      po.template run<decltype(lv)::value>(
          [&](unsigned long long e) {
            return at(sm.S(), unsigned(e) & 0xffffu) * at(sm.S(), unsigned(e >> 16) & 0xffffu) * at(sm.S(), unsigned(e >> 32) & 0xffffu);
          },
          [&](int t, int flag, double acc) {
            const double v = sm.S()[t] - acc;
            sm.S()[t] = flag ? inv_pivot(t - DG0, v) : v;
          });
      sync_phase();
    });
#ifdef CPG_IPM_HOST_EMU
    phase([&](int tid) { tail_factor(tid); });
#else
    tail_factor(int(threadIdx.x));
    __syncthreads();
#endif
  }

  // ---- triangular solves, in place on u (k-space)
  // value of right-hand-side entry k as ldl_solve expects it: leaves of the elimination tree pre-multiplied by 1/d_k
  IPM_FN double lead(int k, double b) const { return ((sm.l0mask()[k >> 4] >> (k & 15)) & 1) ? b * sm.S()[DG0 + k] : b; }
  IPM_FN void tail_solve(int tid, double* u);
  IPM_NOINLINE void ldl_solve(double* u) {
    IPM_COUNT(0);
    const PlanFw pf = plan_fw();
    const PlanBw pb = plan_bw();
    // forward: row k of a wide level becomes D^-1 L^-1 b at once (it is final when its level is done); the tail rows
    // (flag) only collect the contributions of the wide columns.  The leaves (level 0) arrive already scaled: see lead().
    static_for<1, NLW + 1>([&](auto lv) {
      pf.template run<decltype(lv)::value>(
          [&](uint32_t e) { return at(sm.S(), e & 0xffffu) * at(u, e >> 16); },
          [&](int k, int tail, double acc) { const double v = u[k] - acc; u[k] = tail ? v : v * sm.S()[DG0 + k]; });
      sync_phase();
    });
    phase([&](int tid) { tail_solve(tid, u); });
    static_for<0, NLW>([&](auto lv) {
      pb.template run<decltype(lv)::value>(
          [&](uint32_t e) { return at(sm.S(), e & 0xffffu) * at(u, e >> 16); },
          [&](int k, int, double acc) { u[k] -= sm.S()[DG0 + k] * acc; });
      sync_phase();
    });
  }

  // ---- KKT solve with iterative refinement (kkt_solve, kkt.c:87-265); rhs(k) -> out (k-space); returns #refinements
  // right-hand sides of the KKT systems: the two initialisation solves (ecos.c:300-420), the constant one (RHS1) and
  // the vector assembled in sm.rhs (affine / combined direction)
  enum { RHS_INIT_P = 0, RHS_INIT_D, RHS_ONE, RHS_VEC };
  IPM_FN double rhs_of(int mode, int k) const {
    if (mode == RHS_VEC) return sm.rhs()[k];
    const double c = sm.cbh()[k];
    if (mode == RHS_ONE) return k < N ? -c : c;
    if (mode == RHS_INIT_P) return k < N ? 0.0 : c;
    return k < N ? -c : 0.0;
  }
  IPM_NOINLINE int kkt_solve(int mode, double* out, bool isinit) {
    IPM_COUNT(3);
    PerThread<PT> dpx;
    auto rhs = [&](int k) { return rhs_of(mode, k); };
    double s_[1], m_[1];
    phase_red<0, 1>(sm, rb, s_, m_, [&](int tid, double*, double* m) {
      each_k(tid, NK, [&](int k) { const double b = rhs(k); out[k] = lead(k, b); m[0] = fmax(m[0], fabs(b)); });
    });
    const double thr = (1.0 + m_[0]) * kLinsysAcc;
    ldl_solve(out);
    double nerr_prev = NAN;
    int kref = 0;
    for (;;) {
      // error e = b - K out (with the static regularisation written exactly like the reference does): rows outside the
      // second-order cones are completed by the owner of the row in the plan; cone rows get the constant part of K here
      // and the scaling block from the cone warps (sm.ce()), the two are added in the reduction below
      double mloc = 0.0;                    // largest |e| among the rows this thread completes
      plan_mv().run<0>([&](uint32_t en) { return at(sm.ag(), en & 0xffffu) * at(out, en >> 16); },
                    [&](int k, int, double acc) {
                      double e = -acc;
                      if (k < ZOFF + L) {
                        const double o = out[k];
                        double b = rhs(k);
                        if (k < N) b -= kDeltaStat * o;
                        else if (k < ZOFF) b += kDeltaStat * o;
                        else b += kDeltaStat * o + (isinit ? o : sm.v()[k - ZOFF] * o);
                        e += b;
                        mloc = fmax(mloc, fabs(e));
                        e = lead(k, e);
                      }
                      sm.e()[k] = e;
                    });
      cones_then_sync([&](int c, const WarpOps& W) {
        const int so = ZOFF + kSocSo[c], d = kSocD[c];
        double* ce = sm.ce() + (kSocSo[c] - L);
        if (isinit) {
          W.each(0, d + 2, [&](int r) {
            const int k = so + r; const double o = out[k];
            ce[r] = r < d ? rhs(k) + (r < d - 1 ? kDeltaStat : -kDeltaStat) * o + o : o;
          });
        } else {
          const double* q = sm.q() + kSocQo[c]; const double* sc = sm.sc() + 8 * c;
          const double e2 = sc[SC_ETA2], d1 = sc[SC_D1], u0 = sc[SC_U0], u1 = sc[SC_U1], v1 = sc[SC_V1];
          const double x1 = out[so], x3 = out[so + d], x4 = out[so + d + 1];
          const double qtx2 = W.sum(0, d - 1, [&](int i) { return q[i] * out[so + 1 + i]; });
          const double vu = v1 * x3 + u1 * x4;
          W.each(0, d - 1, [&](int i) {
            const int k = so + 1 + i; const double o = out[k];
            ce[1 + i] = rhs(k) + (i + 1 < d - 1 ? kDeltaStat : -kDeltaStat) * o + e2 * (o + vu * q[i]);
          });
          if (W.leader()) {
            ce[0] = rhs(so) + (d > 1 ? kDeltaStat : -kDeltaStat) * x1 + e2 * (d1 * x1 + u0 * x4);
            ce[d] = e2 * (v1 * qtx2 + x3);
            ce[d + 1] = e2 * (u0 * x1 + u1 * qtx2 - x4);
          }
        }
      });
      phase_red<0, 1>(sm, rb, s_, m_, [&](int tid, double*, double* m) {
        m[0] = fmax(m[0], mloc);
        each_k(tid, NCR, [&](int i) {
          const int k = ZOFF + L + i;
          const double v = sm.e()[k] + sm.ce()[i];
          sm.e()[k] = lead(k, v);
          m[0] = fmax(m[0], fabs(v));
        });
      });
      const double nerr = m_[0];
      if (kref > 0 && nerr > nerr_prev) {                 // refinement made it worse: undo and stop
        phase([&](int tid) {
#pragma unroll
          for (int j = 0; j < PT; ++j) { const int k = tid + j * T; if (k < NK) out[k] -= dpx.at(tid, j); }
        });
        --kref;
        break;
      }
      if (kref == kNitref || nerr < thr || (kref > 0 && nerr_prev < kIrErrFact * nerr)) break;
      nerr_prev = nerr;
      ldl_solve(sm.e());
      phase([&](int tid) {
#pragma unroll
        for (int j = 0; j < PT; ++j) { const int k = tid + j * T; if (k < NK) { const double dv = sm.e()[k]; dpx.at(tid, j) = dv; out[k] += dv; } }
      });
      ++kref;
    }
    return kref;
  }

  // ---- cone helpers (one warp per cone)
  // lambda = W v for the cone: out may alias v
  IPM_FN void cone_scale(int c, const WarpOps& W, const double* v, double* out) const {
    const int so = kSocSo[c], d = kSocD[c];
    const double* q = sm.q() + kSocQo[c]; const double* sc = sm.sc() + 8 * c;
    const double zeta = W.sum(0, d - 1, [&](int i) { return q[i] * v[so + 1 + i]; });
    const double v0 = v[so];
    const double factor = v0 + safediv(zeta, 1.0 + sc[SC_A]);
    const double eta = sc[SC_ETA];
    W.sync();
    W.each(0, d - 1, [&](int i) { out[so + 1 + i] = eta * (v[so + 1 + i] + factor * q[i]); });
    if (W.leader()) out[so] = eta * (sc[SC_A] * v0 + zeta);
    W.sync();
  }
  // line-search contribution of one cone (ecos.c:985-1035): returns the conic step (0 = no restriction)
  IPM_FN double cone_step(int c, const WarpOps& W, const double* ds, const double* dz) const {
    const int so = kSocSo[c], d = kSocD[c];
    const double* lk = sm.lam() + so;
    const double l0 = lk[0];
    const double n2 = l0 * l0 - W.sum(1, d, [&](int j) { return lk[j] * lk[j]; });
    if (n2 <= 0.0) return 0.0;
    const double lkn = sqrt(n2), inv = 1.0 / lkn;
    const double lb0 = l0 / lkn;
    double step = 0.0;
    for (int which = 0; which < 2; ++which) {
      const double* dv = (which == 0 ? ds : dz) + so;
      const double lt = lb0 * dv[0] - W.sum(1, d, [&](int j) { return (lk[j] / lkn) * dv[j]; });
      const double r0 = inv * lt;
      const double factor = (lt + dv[0]) / (lb0 + 1.0);
      const double nr = sqrt(W.sum(1, d, [&](int j) { const double r = inv * (dv[j] - factor * (lk[j] / lkn)); return r * r; })) - r0;
      if (nr > step) step = nr;
    }
    return step;
  }

  IPM_FN int solve_instance(int inst, const IpmIO& io, double* best);
};

// ----------------------------------------------------------------------------------------------------------------------
// dense tail block: one warp.  tail_factor leaves the inverse pivots in the diagonal slots and X = inv(L_tail) (strictly
// lower part) in the block, so that a solve with the block is two small dense products instead of two dependent sweeps.
IPM_FN void Solver::tail_factor(int tid) {
#ifdef CPG_IPM_HOST_EMU
  if (tid != 0 || NT == 0) return;
  double* B = sm.S() + TT0;
  double dv[NT > 0 ? NT : 1];
  for (int j = 0; j < NT; ++j) {
    const int kj = sm.tail_k()[j];
    const double ij = inv_pivot(kj, sm.S()[DG0 + kj]);
    dv[j] = ij;
    for (int i = j + 1; i < NT; ++i) {
      const double sij = B[i * NT + j];
      for (int k = j + 1; k < i; ++k) B[i * NT + k] -= sij * B[k * NT + j] * ij;
      sm.S()[DG0 + sm.tail_k()[i]] -= sij * sij * ij;
    }
  }
  double X[(NT > 0 ? NT : 1) * (NT > 0 ? NT : 1)];
  for (int c = 0; c < NT; ++c)
    for (int m = 0; m < NT; ++m) {
      double a = 0.0;
      for (int k = 0; k < m; ++k) a -= (B[m * NT + k] * dv[k]) * X[k * NT + c];
      X[m * NT + c] = m == c ? 1.0 : a;
    }
  for (int j = 0; j < NT; ++j) sm.S()[DG0 + sm.tail_k()[j]] = dv[j];
  for (int m = 0; m < NT; ++m) for (int c = 0; c < m; ++c) B[m * NT + c] = X[m * NT + c];
#else
  // Right-looking elimination with one thread per entry (i, k), k <= i, of the block: step j takes the pivot (every
  // participating thread inverts it itself -- the inverse is kept in sm.tw so that the raw pivot stays readable), updates
  // the trailing entries, one CTA barrier per step.  Then one warp turns the unit triangle into its inverse, lane = column.
  constexpr int NTT = NT > 0 ? NT : 1;
  constexpr int NEL = NT * (NT + 1) / 2;
  static_assert(NEL <= T, "one thread per entry of the tail block");
  if (NT == 0) return;
  double* B = sm.S() + TT0;
  int ei = 0, ek = 0;                                   // (i, k) of this thread's entry: tid = i (i + 1) / 2 + k
  if (tid < NEL) { while ((ei + 1) * (ei + 2) / 2 <= tid) ++ei; ek = tid - ei * (ei + 1) / 2; }
  const int kslot = tid < NEL ? (ei == ek ? DG0 + int(sm.tail_k()[ei]) : TT0 + ei * NT + ek) : 0;
  for (int j = 0; j < NT; ++j) {
    if (tid < NEL && ek >= j) {
      const int kj = sm.tail_k()[j];
      const double ij = inv_pivot(kj, sm.S()[DG0 + kj]);
      if (ek > j) sm.S()[kslot] -= (B[ei * NT + j] * B[ek * NT + j]) * ij;
      else if (ei == j) sm.tw()[j] = ij;                 // the thread of entry (j, j)
    }
    __syncthreads();
  }
  if (tid >= 32) return;
  const int c = tid;                                    // column of inv(L) held by this lane
  double x[NTT];
#pragma unroll
  for (int m = 0; m < NT; ++m) {
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < m; ++k) a -= (B[m * NT + k] * sm.tw()[k]) * x[k];
    x[m] = m == c ? 1.0 : a;
  }
  __syncwarp();
#pragma unroll
  for (int m = 1; m < NT; ++m) if (c < m && c < NT) B[m * NT + c] = x[m];
  if (c < NT) sm.S()[DG0 + sm.tail_k()[c]] = sm.tw()[c];
#endif
}

IPM_FN void Solver::tail_solve(int tid, double* u) {
#ifdef CPG_IPM_HOST_EMU
  if (tid != 0 || NT == 0) return;
  const double* X = sm.S() + TT0;
  double b[NT > 0 ? NT : 1], w[NT > 0 ? NT : 1];
  for (int i = 0; i < NT; ++i) b[i] = u[sm.tail_k()[i]];
  for (int i = 0; i < NT; ++i) {
    double y = b[i];
    for (int j = 0; j < i; ++j) y += X[i * NT + j] * b[j];
    w[i] = y * sm.S()[DG0 + sm.tail_k()[i]];
  }
  for (int i = 0; i < NT; ++i) {
    double xv = w[i];
    for (int m = i + 1; m < NT; ++m) xv += X[m * NT + i] * w[m];
    u[sm.tail_k()[i]] = xv;
  }
#else
  if (NT == 0 || tid >= 32) return;
  const int i = tid;
  const bool on = i < NT;
  const int ki = on ? sm.tail_k()[i] : 0;
  const double* X = sm.S() + TT0;
  double y = on ? u[ki] : 0.0;
  sm.tw()[i] = y;
  __syncwarp();
#pragma unroll
  for (int j = 0; j < NT - 1; ++j) if (on && j < i) y += X[i * NT + j] * sm.tw()[j];
  const double w = on ? y * sm.S()[DG0 + ki] : 0.0;
  sm.tw()[32 + i] = w;
  __syncwarp();
  double xv = w;
#pragma unroll
  for (int m = 1; m < NT; ++m) if (on && m > i) xv += X[m * NT + i] * sm.tw()[32 + m];
  if (on) u[ki] = xv;
#endif
}

// ----------------------------------------------------------------------------------------------------------------------
struct Stats { double gap, mu, kapovert, pcost, dcost, relgap, pres, dres, pinfres, dinfres; };

IPM_FN bool better(const Stats& a, const Stats& b) {        // compareStatistics, ecos.c:61-98
  const bool g = a.gap > 0 && b.gap > 0 && a.gap < b.gap;
  const bool mu = a.mu > 0 && a.mu < b.mu;
  if (a.kapovert > 1) return g && (a.pinfres > 0 && a.pinfres < b.pres) && mu;
  return g && (a.pres > 0 && a.pres < b.pres) && (a.dres > 0 && a.dres < b.dres) &&
         (a.kapovert > 0 && a.kapovert < b.kapovert) && mu;
}

IPM_FN int Solver::solve_instance(int inst, const IpmIO& io, double* best) {
  double s_[16], m_[4];
  const double* par = io.params + size_t(inst) * NPB;
  double* const zv = sm.xyz() + ZOFF;           // z (stretched)
  double* const wdz = sm.e() + ZOFF;            // W dz lives in the error vector between solves

  // ---- cpg_canonicalize: c, b, h from the user parameters (equilibrated), cvxpygen/utils.py:279-294 + equil.c:326-338
  phase([&](int tid) { each_k(tid, NK, [&](int k) { sm.cbh()[k] = gm.cbh_base[k]; }); });
  phase([&](int tid) { each_k(tid, NMAP, [&](int e) { atomic_add(sm.cbh() + gm.map_t[e], gm.map_v[e] * par[gm.map_p[e]]); }); });
#if IPM_MATPAR
  // ---- a user parameter enters G or A: this instance's matrix values, then what ECOS_updateData does with new values
  // (ecos.c:1648-1695) -- set_equilibration from scratch (equil.c:210-342): EQUIL_ITERS passes of  scale = sqrt(max |entry|)  per row and
  // per column of [A ; G] (one value, the SUM of the row maxima, for all rows of a second-order cone; values below 1e-6 -> 1), rows
  // divided first, then columns; c, b, h divided by the accumulated scalings.  The maxima are exact and order-independent (atomic max
  // on the bit patterns of non-negative doubles); the cone sum runs in row order on one thread like the reference's loop.
  // Work vectors: rhs = this pass's scaling per k-row, px = accumulated scaling.
  double* const usc = best + NK + MT;
  phase([&](int tid) {
    each_k(tid, NNZM, [&](int e) { sm.ag()[e] = gm.ent_base[e]; });
    each_k(tid, NK, [&](int k) { sm.px()[k] = 1.0; });
  });
  phase([&](int tid) { each_k(tid, IPM_NEMAP, [&](int i) { atomic_add(sm.ag() + gm.emap_t[i], gm.emap_v[i] * par[gm.emap_p[i]]); }); });
  for (int pass = 0; pass < 3; ++pass) {
    phase([&](int tid) { each_k(tid, NK, [&](int k) { sm.rhs()[k] = 0.0; }); });
    phase([&](int tid) {
      each_k(tid, NNZM, [&](int e) {
        const double v = fabs(sm.ag()[e]);
        atomic_max_nonneg(sm.rhs() + gm.mr_t[e], v); atomic_max_nonneg(sm.rhs() + gm.mr_s[e], v);
      });
    });
    phase([&](int tid) {
      if (tid < NSOC) {
        const int so = ZOFF + kSocSo[tid], d = kSocD[tid];
        double tot = 0.0;
        for (int r = 0; r < d; ++r) tot += sm.rhs()[so + r];
        for (int r = 0; r < d; ++r) sm.rhs()[so + r] = tot;
      }
    });
    phase([&](int tid) {
      each_k(tid, NK, [&](int k) {
        const double v = sm.rhs()[k], f = fabs(v) < 1e-6 ? 1.0 : sqrt(v);
        sm.rhs()[k] = f; sm.px()[k] *= f;
      });
    });
    phase([&](int tid) { each_k(tid, NNZM, [&](int e) { sm.ag()[e] = (sm.ag()[e] / sm.rhs()[gm.mr_t[e]]) / sm.rhs()[gm.mr_s[e]]; }); });
  }
  phase([&](int tid) { each_k(tid, NK, [&](int k) { const double f = sm.px()[k]; sm.cbh()[k] /= f; usc[k] = 1.0 / f; }); });
#else
  const double* const usc = gm.unscale;
#endif
  phase_red<3, 0>(sm, rb, s_, m_, [&](int tid, double* s, double*) {
    each_k(tid, NK, [&](int k) { const double v = sm.cbh()[k]; s[k < N ? 0 : (k < ZOFF ? 1 : 2)] += v * v; });
  });
  const double resx0 = fmax(1.0, sqrt(s_[0])), resy0 = fmax(1.0, sqrt(s_[1])), resz0 = fmax(1.0, sqrt(s_[2]));

  // ---- init (ecos.c:260-452): K with the identity scaling, two least-squares solves, bring2cone
  phase([&](int tid) {
    each_k(tid, NS, [&](int i) {
      double v = gm.Sbase[i];
      if (i >= DG0 + ZOFF && i < DG0 + NK) v += sign_of(i - DG0) > 0 ? 1.0 : -1.0;
      sm.S()[i] = v;
    });
    each_k(tid, MT, [&](int i) { sm.sv()[i] = 0.0; sm.lam()[i] = 0.0; sm.rz()[i] = 0.0; sm.dsw()[i] = 0.0; });
    each_k(tid, NK, [&](int k) { sm.xyz()[k] = 0.0; });
  });
#if IPM_MATPAR
  phase([&](int tid) { each_k(tid, NNZM, [&](int e) { sm.S()[gm.ag_slot[e]] = sm.ag()[e]; }); });
#endif
  factor();
  auto bring2cone = [&](const double* r, double rsign, double* out) {        // out = rsign * r shifted into the cone
    phase_red_cones<0, 1>(sm, rb, s_, m_,
        [&](int tid, double*, double* m) { each_k(tid, L, [&](int i) { const double v = rsign * r[i]; if (v <= 0) m[0] = fmax(m[0], -v); }); },
        [&](int c, const WarpOps& W) {
          const int so = kSocSo[c], d = kSocD[c];
          const double nr = sqrt(W.sum(1, d, [&](int j) { return r[so + j] * r[so + j]; }));
          if (W.leader()) sm.cone()[c] = rsign * r[so] - nr;
        });
    double alpha = fmax(-kGamma, m_[0]);
    for (int c = 0; c < NSOC; ++c) { const double cres = sm.cone()[c]; if (cres <= 0 && -cres > alpha) alpha = -cres; }
    alpha += 1.0;
    phase_cones([&](int tid) { each_k(tid, L, [&](int i) { out[i] = rsign * r[i] + alpha; }); },
                [&](int c, const WarpOps& W) {
                  const int so = kSocSo[c], d = kSocD[c];
                  W.each(0, d, [&](int j) { out[so + j] = rsign * r[so + j] + (j == 0 ? alpha : 0.0); });
                });
  };
  kkt_solve(RHS_INIT_P, sm.px(), true);
  phase([&](int tid) { each_k(tid, N, [&](int k) { sm.xyz()[k] = sm.px()[k]; }); });
  bring2cone(sm.px() + ZOFF, -1.0, sm.sv());
  kkt_solve(RHS_INIT_D, sm.px(), true);
  phase([&](int tid) { each_k(tid, P, [&](int i) { sm.xyz()[N + i] = sm.px()[N + i]; }); });
  bring2cone(sm.px() + ZOFF, 1.0, zv);
  double kap = 1.0, tau = 1.0;

  // ---- main loop (ecos.c:1123-1583)
  Stats I{}, Ibest{};
  double cx = 0, by = 0, hz = 0, rt = 0, best_kap = 1, best_tau = 1, best_cx = 0, best_by = 0, best_hz = 0;
  double pres_prev = NAN, step = 0.0;
  int exitcode = kFatal, it = 0;
  auto save_best = [&]() {
    Ibest = I; best_kap = kap; best_tau = tau; best_cx = cx; best_by = by; best_hz = hz;
    phase([&](int tid) { each_k(tid, NK, [&](int k) { best[k] = sm.xyz()[k]; }); each_k(tid, MT, [&](int i) { best[NK + i] = sm.sv()[i]; }); });
  };
  auto restore_best = [&]() {
    I = Ibest; kap = best_kap; tau = best_tau; cx = best_cx; by = best_by; hz = best_hz;
    phase([&](int tid) { each_k(tid, NK, [&](int k) { sm.xyz()[k] = best[k]; }); each_k(tid, MT, [&](int i) { sm.sv()[i] = best[NK + i]; }); });
  };
  auto check_exit = [&](int mode) {                          // checkExitConditions, ecos.c:179-257
    const double feastol = mode ? stg.feastol_inacc : stg.feastol, abstol = mode ? stg.abstol_inacc : stg.abstol,
                 reltol = mode ? stg.reltol_inacc : stg.reltol;
    if ((-cx > 0 || -by - hz >= -abstol) && (I.pres < feastol && I.dres < feastol) && (I.gap < abstol || I.relgap < reltol))
      return int(kOptimal) + mode;
    if (I.dinfres < feastol && tau < kap) return int(kDinf) + mode;
    if ((I.pinfres < feastol && tau < kap) || (tau < stg.feastol && kap < stg.feastol && I.pinfres < stg.feastol))
      return int(kPinf) + mode;
    return int(kNotConverged);
  };
  for (;; ++it) {
    // computeResiduals (ecos.c:455-499): -A'y - G'z -> rhs[0:N], A x -> rhs[N:ZOFF], s + G x -> rz
    plan_mv().run<0>([&](uint32_t en) { return at(sm.ag(), en & 0xffffu) * at(sm.xyz(), en >> 16); },
                  [&](int k, int, double acc) {
                    if (k < N) sm.rhs()[k] = -acc;
                    else if (k < ZOFF) sm.rhs()[k] = acc;
                    else sm.rz()[k - ZOFF] = sm.sv()[k - ZOFF] + acc;
                  });
    sync_phase();
    phase_red<16, 0>(sm, rb, s_, m_, [&](int tid, double* s, double*) {
      each_k(tid, N, [&](int k) {
        const double hr = sm.rhs()[k], c = sm.cbh()[k], x = sm.xyz()[k];
        s[0] += hr * hr; s[3] += c * x; s[6] += x * x;
        const double r = hr - tau * c; sm.rhs()[k] = r; s[11] += r * r;
      });
      each_k(tid, P, [&](int i) {
        const int k = N + i; const double hr = sm.rhs()[k], b = sm.cbh()[k], y = sm.xyz()[k];
        s[1] += hr * hr; s[4] += b * y; s[7] += y * y;
        const double r = hr - tau * b; sm.rhs()[k] = r; s[12] += r * r;
      });
      each_k(tid, NS, [&](int i) { sm.S()[i] = gm.Sbase[i]; });      // constant part of K for this iteration's kkt_update
      each_k(tid, MT, [&](int i) {
        const double hr = sm.rz()[i], h = sm.cbh()[ZOFF + i], z = zv[i], sv = sm.sv()[i];
        s[2] += hr * hr; s[5] += h * z; s[8] += sv * sv; s[9] += z * z; s[10] += sv * z;
        const double r = hr - tau * h; sm.rz()[i] = r; s[13] += r * r;
      });
    });
    {
      const double hresx = sqrt(s_[0]), hresy = sqrt(s_[1]), hresz = sqrt(s_[2]);
      cx = s_[3]; by = s_[4]; hz = s_[5];
      const double nx = sqrt(s_[6]), ny = sqrt(s_[7]), ns = sqrt(s_[8]), nz = sqrt(s_[9]);
      rt = kap + cx + by + hz;
      // updateStatistics
      I.gap = s_[10];
      I.mu = (I.gap + kap * tau) / (CONE_D + 1);
      I.kapovert = kap / tau;
      I.pcost = cx / tau; I.dcost = -(hz + by) / tau;
      I.relgap = I.pcost < 0 ? I.gap / (-I.pcost) : (I.dcost > 0 ? I.gap / I.dcost : NAN);
      const double nry = P > 0 ? sqrt(s_[12]) / fmax(resy0 + nx, 1.0) : 0.0;
      const double nrz = sqrt(s_[13]) / fmax(resz0 + nx + ns, 1.0);
      I.pres = fmax(nry, nrz) / tau;
      I.dres = sqrt(s_[11]) / fmax(resx0 + ny + nz, 1.0) / tau;
      I.pinfres = (hz + by) / fmax(ny + nz, 1.0) < -stg.reltol ? hresx / fmax(ny + nz, 1.0) : NAN;
      I.dinfres = cx / fmax(nx, 1.0) < -stg.reltol ? fmax(hresy / fmax(nx, 1.0), hresz / fmax(nx + ns, 1.0)) : NAN;
    }
    // safeguards and exit tests
    if (it > 0 && (I.pres > kSafeguard * pres_prev || I.gap < 0)) {
      restore_best(); exitcode = check_exit(kInaccOffset);
      if (exitcode == kNotConverged) exitcode = kNumerics;
      break;
    }
    pres_prev = I.pres;
    exitcode = check_exit(0);
    if (exitcode != kNotConverged) break;
    if (it > 0 && step == kStepMin * kGamma) {
      restore_best(); exitcode = check_exit(kInaccOffset);
      if (exitcode == kNotConverged) exitcode = kNumerics;
      break;
    }
    if (it == stg.maxit) {
      if (!better(I, Ibest)) restore_best();
      exitcode = check_exit(kInaccOffset);
      if (exitcode == kNotConverged) exitcode = kMaxit;
      break;
    }
    if (isnan(I.pcost)) {
      if (!better(I, Ibest)) restore_best();
      exitcode = check_exit(kInaccOffset);
      if (exitcode == kNotConverged) exitcode = kNumerics;
      break;
    }
    if (it == 0 || better(I, Ibest)) save_best();

    // updateScalings + lambda = W z, and the constant part of the KKT image
    phase_cones(
        [&](int tid) {
          if (tid == 0) *sm.flag() = 0;
#if IPM_MATPAR
          each_k(tid, NNZM, [&](int e) { sm.S()[gm.ag_slot[e]] = sm.ag()[e]; });      // this instance's A, G entries of K (S <- Sbase is done)
#endif
          each_k_nc(tid, L, [&](int i) {
            const double v = safediv(sm.sv()[i], zv[i]), w = sqrt(v);
            sm.v()[i] = v; sm.w()[i] = w; sm.lam()[i] = w * zv[i];
          });
        },
        [&](int c, const WarpOps& W) {
          const int so = kSocSo[c], d = kSocD[c];
          const double* sk = sm.sv() + so; const double* zk = zv + so;
          double* q = sm.q() + kSocQo[c]; double* sc = sm.sc() + 8 * c;
          const double sres = sk[0] * sk[0] - W.sum(1, d, [&](int j) { return sk[j] * sk[j]; });
          const double zres = zk[0] * zk[0] - W.sum(1, d, [&](int j) { return zk[j] * zk[j]; });
          if (sres <= 0 || zres <= 0) { if (W.leader()) sm.cone()[c] = 1.0; return; }
          const double snorm = sqrt(sres), znorm = sqrt(zres);
          const double eta2 = safediv(snorm, znorm), eta = sqrt(eta2);
          const double gamma = sqrt(0.5 * (1.0 + W.sum(0, d, [&](int j) { return safediv(sk[j], snorm) * safediv(zk[j], znorm); })));
          const double o2g = safediv(0.5, gamma);
          const double a = o2g * (safediv(sk[0], snorm) + safediv(zk[0], znorm));
          W.each(1, d, [&](int j) { q[j - 1] = o2g * (safediv(sk[j], snorm) - safediv(zk[j], znorm)); });
          W.sync();
          const double w = W.sum(0, d - 1, [&](int i) { return q[i] * q[i]; });
          const double temp = 1.0 + a;
          const double cc = 1.0 + a + safediv(w, temp);
          const double dd = 1.0 + safediv(2.0, temp) + safediv(w, temp * temp);
          double d1 = 0.5 * (a * a + w * (1.0 - safediv(cc * cc, 1.0 + w * dd)));
          if (d1 < 0) d1 = 0;
          const double u0sq = a * a + w - d1, u0 = sqrt(u0sq);
          const double c2byu02 = safediv(cc * cc, u0sq);
          if (c2byu02 - dd <= 0) { if (W.leader()) sm.cone()[c] = 1.0; return; }
          if (W.leader()) {
            sc[SC_ETA2] = eta2; sc[SC_ETA] = eta; sc[SC_A] = a; sc[SC_D1] = d1; sc[SC_U0] = u0;
            sc[SC_U1] = sqrt(c2byu02); sc[SC_V1] = sqrt(c2byu02 - dd); sc[SC_W] = w;
            sm.cone()[c] = 0.0;
          }
          W.sync();
          cone_scale(c, W, zv, sm.lam());
        });
    {
      bool out = false;
      for (int c = 0; c < NSOC; ++c) out = out || sm.cone()[c] != 0.0;
      if (out) {
        restore_best(); exitcode = check_exit(kInaccOffset);
        if (exitcode == kNotConverged) exitcode = kOutcone;
        break;
      }
    }
    // kkt_update: the scaling block
    phase_cones(
        [&](int tid) { each_k_nc(tid, L, [&](int i) { sm.S()[DG0 + ZOFF + i] = -sm.v()[i] - kDeltaStat; }); },
        [&](int c, const WarpOps& W) {
          const int so = kSocSo[c], d = kSocD[c];
          const double* q = sm.q() + kSocQo[c]; const double* sc = sm.sc() + 8 * c;
          const double e2 = sc[SC_ETA2];
          const uint16_t* sv_ = sm.socv() + kSocVo[c]; const uint16_t* su_ = sm.socu() + kSocUo[c];
          W.each(0, d, [&](int r) { sm.S()[DG0 + ZOFF + so + r] = (r == 0 ? -e2 * sc[SC_D1] : -e2) - kDeltaStat; });
          W.each(0, d - 1, [&](int i) { sm.S()[sv_[i]] = -e2 * sc[SC_V1] * q[i]; sm.S()[su_[1 + i]] = -e2 * sc[SC_U1] * q[i]; });
          if (W.leader()) {
            sm.S()[su_[0]] = -e2 * sc[SC_U0];
            sm.S()[DG0 + ZOFF + so + d] = -e2;
            sm.S()[DG0 + ZOFF + so + d + 1] = e2 + kDeltaStat;
          }
        });
    factor();
    // search directions
    kkt_solve(RHS_ONE, sm.sol1(), false);
    phase([&](int tid) {                                      // RHS_affine
      each_k(tid, P, [&](int i) { sm.rhs()[N + i] = -sm.rhs()[N + i]; });
      each_k(tid, MT, [&](int i) { sm.rhs()[ZOFF + i] = sm.sv()[i] - sm.rz()[i]; });
    });
    kkt_solve(RHS_VEC, sm.px(), false);
    phase_red<2, 0>(sm, rb, s_, m_, [&](int tid, double* s, double*) {
      each_k(tid, NK, [&](int k) { const double c = sm.cbh()[k]; s[0] += c * sm.sol1()[k]; s[1] += c * sm.px()[k]; });
    });
    const double dtau_denom = kap / tau - s_[0];
    const double dtauaff = (rt - kap + s_[1]) / dtau_denom;
    const double dkapaff = -kap - kap / tau * dtauaff;
    // dzaff, W dzaff, W\dsaff and the affine line search
    auto direction_and_linesearch = [&](double dt, bool combined, double dtau_, double dkap_) -> double {
      // px_z += dt * sol1_z (all of px when combined); wdz = W px_z; dsw = -(combined ? dsw : lam) - wdz
      phase_red_cones<0, 2>(sm, rb, s_, m_,
          [&](int tid, double*, double* m) {
            each_k_nc(tid, combined ? ZOFF : 0, [&](int k) { sm.px()[k] += dt * sm.sol1()[k]; });
            each_k_nc(tid, L, [&](int i) {
              const double dz = sm.px()[ZOFF + i] + dt * sm.sol1()[ZOFF + i];
              sm.px()[ZOFF + i] = dz;
              const double wz = sm.w()[i] * dz, lam = sm.lam()[i];
              const double dsw = -(combined ? sm.dsw()[i] : lam) - wz;
              wdz[i] = wz; sm.dsw()[i] = dsw;
              m[0] = fmax(m[0], -dsw / lam); m[1] = fmax(m[1], -wz / lam);
            });
          },
          [&](int c, const WarpOps& W) {
            const int so = kSocSo[c], d = kSocD[c];
            W.each(0, d + 2, [&](int r) { sm.px()[ZOFF + so + r] += dt * sm.sol1()[ZOFF + so + r]; });
            W.sync();
            cone_scale(c, W, sm.px() + ZOFF, wdz);
            W.each(0, d, [&](int r) { sm.dsw()[so + r] = -(combined ? sm.dsw()[so + r] : sm.lam()[so + r]) - wdz[so + r]; });
            W.sync();
            const double st = cone_step(c, W, sm.dsw(), wdz);
            if (W.leader()) sm.cone()[c] = st;
          });
      // lineSearch, ecos.c:947-1046 (m_[0] = -rhomin, m_[1] = -sigmamin)
      double alpha;
      if (L > 0) {
        const double rhomin = -m_[0], sigmamin = -m_[1];
        if (-sigmamin > -rhomin) alpha = sigmamin < 0 ? 1.0 / (-sigmamin) : 1.0 / kEps;
        else alpha = rhomin < 0 ? 1.0 / (-rhomin) : 1.0 / kEps;
      } else alpha = 10.0;
      const double mtt = -tau / dtau_, mkk = -kap / dkap_;
      if (mtt > 0 && mtt < alpha) alpha = mtt;
      if (mkk > 0 && mkk < alpha) alpha = mkk;
      for (int c = 0; c < NSOC; ++c) { const double st = sm.cone()[c]; if (st != 0.0) { const double t_ = 1.0 / st; if (t_ < alpha) alpha = t_; } }
      if (alpha > kStepMax) alpha = kStepMax;
      if (alpha < kStepMin) alpha = kStepMin;
      return alpha;
    };
    const double step_aff = direction_and_linesearch(dtauaff, false, dtauaff, dkapaff);
    double sigma = 1.0 - step_aff; sigma = sigma * sigma * sigma;
    if (sigma > kSigmaMax) sigma = kSigmaMax;
    if (sigma < kSigmaMin) sigma = kSigmaMin;
    const double sigmamu = sigma * I.mu, oms = 1.0 - sigma;
    // RHS_combined (ecos.c:688-757): dsw <- lambda \ (lambda o lambda + dsw o wdz - sigma mu e); rhs_z = -(1-sigma) rz + W dsw
    phase_cones(
        [&](int tid) {
          each_k_nc(tid, ZOFF, [&](int k) { sm.rhs()[k] *= oms; });
          each_k_nc(tid, L, [&](int i) {
            const double lam = sm.lam()[i];
            const double ds1 = lam * lam + sm.dsw()[i] * wdz[i] - sigmamu;
            const double dv = safediv(ds1, lam);
            sm.dsw()[i] = dv;
            sm.rhs()[ZOFF + i] = -oms * sm.rz()[i] + sm.w()[i] * dv;
          });
        },
        [&](int c, const WarpOps& W) {
          const int so = kSocSo[c], d = kSocD[c];
          const double* lk = sm.lam() + so; double* ds = sm.dsw() + so; const double* wz = wdz + so;
          double* tmp = sm.px() + ZOFF + so;                       // px is free between the two solves
          const double l0 = lk[0], ds0 = ds[0], wz0 = wz[0];
          const double ll = W.sum(0, d, [&](int j) { return lk[j] * lk[j]; });
          const double dw = W.sum(0, d, [&](int j) { return ds[j] * wz[j]; });
          W.each(1, d, [&](int j) { tmp[j] = 2.0 * l0 * lk[j] + ds0 * wz[j] + wz0 * ds[j]; });
          const double w0 = ll + dw - sigmamu;
          W.sync();
          // conicDivision(lambda, ds1)
          const double rho = l0 * l0 - W.sum(1, d, [&](int j) { return lk[j] * lk[j]; });
          const double zeta = W.sum(1, d, [&](int j) { return lk[j] * tmp[j]; });
          const double factor = safediv(safediv(zeta, l0) - w0, rho);
          W.each(1, d, [&](int j) { ds[j] = factor * lk[j] + safediv(tmp[j], l0); });
          if (W.leader()) ds[0] = safediv(l0 * w0 - zeta, rho);
          W.sync();
          cone_scale(c, W, sm.dsw(), sm.px() + ZOFF);
          W.each(0, d, [&](int r) { sm.rhs()[ZOFF + so + r] = -oms * sm.rz()[so + r] + tmp[r]; });
          if (W.leader()) { sm.rhs()[ZOFF + so + d] = 0.0; sm.rhs()[ZOFF + so + d + 1] = 0.0; }
          W.sync();
        });
    kkt_solve(RHS_VEC, sm.px(), false);
    phase_red<1, 0>(sm, rb, s_, m_, [&](int tid, double* s, double*) {
      each_k(tid, NK, [&](int k) { s[0] += sm.cbh()[k] * sm.px()[k]; });
    });
    const double bkap = kap * tau + dkapaff * dtauaff - sigmamu;
    const double dtau = (oms * rt - bkap / tau + s_[0]) / dtau_denom;
    const double dkap = -(bkap + kap * dtau) / tau;
    step = direction_and_linesearch(dtau, true, dtau, dkap) * kGamma;
    // ds = W (W\ds); update the iterate
    phase_cones(
        [&](int tid) {
          each_k_nc(tid, ZOFF + L, [&](int k) { sm.xyz()[k] += step * sm.px()[k]; });
          each_k_nc(tid, L, [&](int i) { sm.sv()[i] += step * (sm.w()[i] * sm.dsw()[i]); });
        },
        [&](int c, const WarpOps& W) {
          const int so = kSocSo[c], d = kSocD[c];
          cone_scale(c, W, sm.dsw(), wdz);
          W.each(0, d, [&](int r) { sm.sv()[so + r] += step * wdz[so + r]; zv[so + r] += step * sm.px()[ZOFF + so + r]; });
        });
    kap += step * dkap; tau += step * dtau;
  }

  // ---- backscale (ecos.c:1051-1070) and retrieval (cpg_retrieve_prim / dual / info, cvxpygen/utils.py:950-985)
  phase([&](int tid) {
    const double it_ = 1.0 / tau;
    (void)it_;
    each_k(tid, NPRIM, [&](int i) { const int k = gm.prim_idx[i]; io.prim[size_t(inst) * NPRIM + i] = sm.xyz()[k] * usc[k] / tau; });
    each_k(tid, NDUAL, [&](int i) { const int k = gm.dual_idx[i]; io.dual[size_t(inst) * NDUAL + i] = sm.xyz()[k] * usc[k] / tau; });
    if (io.sol_x) each_k(tid, N, [&](int k) { io.sol_x[size_t(inst) * N + k] = sm.xyz()[k] * usc[k] / tau; });
    if (io.sol_y) each_k(tid, P, [&](int i) { io.sol_y[size_t(inst) * P + i] = sm.xyz()[N + i] * usc[N + i] / tau; });
    if (io.sol_z || io.sol_s) {
      each_k(tid, L, [&](int i) {
        if (io.sol_z) io.sol_z[size_t(inst) * M + i] = zv[i] * usc[ZOFF + i] / tau;
        if (io.sol_s) io.sol_s[size_t(inst) * M + i] = sm.sv()[i] / (usc[ZOFF + i] * tau);
      });
      int o = L;
      for (int c = 0; c < NSOC; ++c) {
        const int so = kSocSo[c], d = kSocD[c];
        each_k(tid, d, [&](int r) {
          if (io.sol_z) io.sol_z[size_t(inst) * M + o + r] = zv[so + r] * usc[ZOFF + so + r] / tau;
          if (io.sol_s) io.sol_s[size_t(inst) * M + o + r] = sm.sv()[so + r] / (usc[ZOFF + so + r] * tau);
        });
        o += d;
      }
    }
    if (tid == 0) {
      const double obj = I.pcost + IPM_D_CONST;
      io.obj_val[inst] = IPM_IS_MAX ? -obj : obj;
      io.iter[inst] = it; io.status[inst] = exitcode; io.pri_res[inst] = I.pres; io.dua_res[inst] = I.dres;
    }
  });
  return exitcode;
}

#ifndef CPG_IPM_HOST_EMU
// ----------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CPG_IPM_THREADS, 1)
ipm_kernel(const unsigned char* __restrict__ smem_blob, const unsigned char* __restrict__ gmem_blob, IpmSettings stg, IpmIO io) {
  __shared__ int next_inst;
  Solver sv;
    sv.gm = make_gm(gmem_blob);
  sv.stg = stg; sv.rb = 0;
  // stage the constant tables: [f64 ag_val, 0 | u32 plan entries | u16 descriptors and index lists] -- the blob is laid out
  // exactly like its shared-memory image (O_AG onwards), so it arrives as TMA bulk copies (cp.async.bulk, 32 KB chunks) that
  // complete on one mbarrier; the null slot of S is set meanwhile
  {
    static_assert(IPM_SB_BYTES % 16 == 0 && (O_AG * 8) % 16 == 0, "bulk copies move 16-byte units");
    static_assert(IPM_SB_U32_OFF == (NNZM + 1) * 8, "the f64 part of the blob is ag_val plus its null entry");
    unsigned char* dst = smem_raw + size_t(O_AG) * 8;
#ifdef CPG_SIMT_HOST_EMU
    if (threadIdx.x == 0) memcpy(dst, smem_blob, IPM_SB_BYTES);
    if (threadIdx.x == 0) sv.sm.S()[NS] = 0.0;
    __syncthreads();
#else
    __shared__ __align__(8) unsigned long long stage_bar;
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&stage_bar);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)IPM_SB_BYTES) : "memory");
      constexpr unsigned CHUNK = 32768;
      for (unsigned off = 0; off < (unsigned)IPM_SB_BYTES; off += CHUNK) {
        const unsigned n = (unsigned)IPM_SB_BYTES - off < CHUNK ? (unsigned)IPM_SB_BYTES - off : CHUNK;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((unsigned)__cvta_generic_to_shared(dst + off)), "l"(smem_blob + off), "r"(n), "r"(bar) : "memory");
      }
      sv.sm.S()[NS] = 0.0;          // the slot every null entry of the plans points at
    }
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(bar), "r"(0) : "memory");
    __syncthreads();
#endif
  }
  double* best = io.best + size_t(blockIdx.x) * BEST_STRIDE;
  for (;;) {
    if (threadIdx.x == 0) next_inst = atomicAdd(io.counter, 1);
    __syncthreads();
    const int inst = next_inst;
    __syncthreads();
    if (inst >= io.B) break;
    sv.solve_instance(inst, io, best);
  }
}
#endif

}  // namespace cpgipm
