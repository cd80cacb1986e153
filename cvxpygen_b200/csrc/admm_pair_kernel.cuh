// admm_pair_kernel.cuh -- main batched ADMM kernel for sm_100a: TWO problem instances per warp.
//
// Why two: ncu on the one-instance-per-warp version showed the kernel bound by shared-memory bandwidth
// (71 % of peak wavefronts; profiles/r1_v1_ncu_summary.md): every inner step of the KKT-solve schedule
// loads one 8-byte coefficient per lane and uses it once.  Register-blocking two instances makes every
// coefficient (and every index) loaded from shared memory feed two FMAs; the two instances' work vectors
// are interleaved (w2[pos] = {w_A[pos], w_B[pos]}, 16 bytes) so one LDS.128 fetches both operands.
// On top of that the chain part of the schedule is encoded as DENSE tiles: all lanes of a row-part read the
// same w2 address (a shared-memory broadcast, one wavefront) and no index is loaded at all
// (offline/schedule.py: encode_sparse / encode_dense).
//
// The two slots of a warp are independent instances: each has its own iteration counter, termination
// check and epilogue; when one terminates its slot is refilled from the global instance queue at once.
// State per slot in registers: x, z, y (lane i%32 owns element i).  q, l, u are NOT kept in registers:
// rows that do not depend on a batched parameter are read from the constants blob (already scaled), rows
// that do are read from a small per-warp table written by the slot's prologue.
//
// Reference functions restated: same list as admm_kernel.cuh (a1-a12); the per-instance arithmetic and its
// order are unchanged, so iterates match the single-instance path bit for bit.
#pragma once
#include "admm_kernel.cuh"

namespace cpgb200 {

struct d2 { double a, b; };     // the two instances of a warp

__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void sts2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

// ---- tile executor on interleaved pairs -------------------------------------------------------------------
// tile table entry (8 ints): kind, f64 offset, aux offset (i32 area), K, r_pad, nrows, rows offset (u16), nseg
__device__ __forceinline__ double2 pair_tile_acc(const int4 h0, const int4 h1, const int* I32, const double* F64,
                                                 const double* w2, int lane) {
  const double* v = F64 + h0.y + lane;
  double ax0 = 0.0, ay0 = 0.0, ax1 = 0.0, ay1 = 0.0;
  if (h0.x == 0) {                       // sparse: gather, two packed column indices per word
    const unsigned* cw = reinterpret_cast<const unsigned*>(I32 + h0.z) + lane;
    const int K2 = h0.w >> 1;
#pragma unroll 2
    for (int k = 0; k < K2; ++k) {
      const unsigned cc = cw[k * LANES];
      const double v0 = v[(2 * k) * LANES], v1 = v[(2 * k + 1) * LANES];
      const double2 p0 = *reinterpret_cast<const double2*>(reinterpret_cast<const char*>(w2) + (cc & 0xffffu));
      const double2 p1 = *reinterpret_cast<const double2*>(reinterpret_cast<const char*>(w2) + (cc >> 16));
      ax0 = fma(v0, p0.x, ax0); ay0 = fma(v0, p0.y, ay0);
      ax1 = fma(v1, p1.x, ax1); ay1 = fma(v1, p1.y, ay1);
    }
  } else {                               // dense: contiguous column segments, broadcast reads, no indices
    const int rp = h1.x, p = LANES / rp, part = lane / rp;
    const int* seg = I32 + h0.z;
    for (int sgi = 0; sgi < h1.w; ++sgi) {
      const int c0 = seg[2 * sgi], Kp = seg[2 * sgi + 1];
      const double* wp = w2 + 2 * (c0 + part);
      int k = 0;
      for (; k + 1 < Kp; k += 2) {
        const double v0 = v[k * LANES], v1 = v[(k + 1) * LANES];
        const double2 p0 = lds2(wp + 2 * p * k);
        const double2 p1 = lds2(wp + 2 * p * (k + 1));
        ax0 = fma(v0, p0.x, ax0); ay0 = fma(v0, p0.y, ay0);
        ax1 = fma(v1, p1.x, ax1); ay1 = fma(v1, p1.y, ay1);
      }
      if (k < Kp) {
        const double v0 = v[k * LANES];
        const double2 p0 = lds2(wp + 2 * p * k);
        ax0 = fma(v0, p0.x, ax0); ay0 = fma(v0, p0.y, ay0);
      }
      v += Kp * LANES;
    }
  }
  double2 acc = make_double2(ax0 + ax1, ay0 + ay1);
  for (int o = 16; o >= h1.x; o >>= 1) {
    acc.x += __shfl_xor_sync(FULL, acc.x, o);
    acc.y += __shfl_xor_sync(FULL, acc.y, o);
  }
  return acc;
}

template <int TRAIL>
__device__ __forceinline__ void pair_kkt_solve(const CpgBlobHeader* H, const int* I32, const double* F64,
                                               const uint16_t* U16, double* w2, int lane) {
  const int4* T = reinterpret_cast<const int4*>(I32 + H->i_tiles2);
  const int nf = H->n_fwd_tiles, nt = H->n_tiles;
  int t = 0;
  for (; t < nf; ++t) {
    const int4 h0 = T[2 * t], h1 = T[2 * t + 1];
    const double2 acc = pair_tile_acc(h0, h1, I32, F64, w2, lane);
    __syncwarp();
    if (lane < h1.y) sts2(w2 + 2 * U16[h1.z + lane], acc.x, acc.y);
    __syncwarp();
  }
  if (TRAIL > 0) {
    double2 tacc[TRAIL > 0 ? TRAIL : 1];
#pragma unroll
    for (int j = 0; j < TRAIL; ++j)
      tacc[j] = (j < H->n_trail_tiles) ? pair_tile_acc(T[2 * (t + j)], T[2 * (t + j) + 1], I32, F64, w2, lane) : make_double2(0.0, 0.0);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < TRAIL; ++j) {
      if (j < H->n_trail_tiles) {
        const int4 h1 = T[2 * (t + j) + 1];
        if (lane < h1.y) sts2(w2 + 2 * U16[h1.z + lane], tacc[j].x, tacc[j].y);
      }
    }
    __syncwarp();
    t += H->n_trail_tiles;
  }
  for (; t < nt; ++t) {
    const int4 h0 = T[2 * t], h1 = T[2 * t + 1];
    const double2 acc = pair_tile_acc(h0, h1, I32, F64, w2, lane);
    __syncwarp();
    if (lane < h1.y) sts2(w2 + 2 * U16[h1.z + lane], acc.x, acc.y);
    __syncwarp();
  }
}

// row-blocked ELL sparse mat-vec on an interleaved pair vector
__device__ __forceinline__ double2 pair_ell_dot(const int* tab, const double* F64, const uint16_t* U16,
                                                const double* vec2, int lane) {
  const int K = tab[0];
  const double* v = F64 + tab[1] + lane;
  const uint16_t* c = U16 + tab[2] + lane;
  double ax = 0.0, ay = 0.0;
  for (int k = 0; k < K; ++k) {
    const double vv = v[k * LANES];
    const double2 p = lds2(vec2 + 2 * c[k * LANES]);
    ax = fma(vv, p.x, ax); ay = fma(vv, p.y, ay);
  }
  return make_double2(ax, ay);
}

__device__ __forceinline__ double sel(const double2 v, int s) { return s ? v.y : v.x; }

}  // namespace cpgb200
#ifdef CPG_FAM_GENERATED_SOLVE
#include "cpg_kkt_solve_gen.cuh"     // straight-line schedule of this family (offline/emit_solve.py)
#endif
namespace cpgb200 {

// ---------------------------------------------------------------- the pair solver
template <class Fam>
__device__ void solve_pairs(const CpgBlobHeader* __restrict__ H, const int* __restrict__ I32,
                            const double* __restrict__ F64, const uint16_t* __restrict__ U16,
                            double* __restrict__ w2, double* __restrict__ bv, const int lane,
                            const BatchIO& io, const Settings& st) {
  constexpr int N = Fam::N, M = Fam::M, NXL = (N + 31) / 32, NZL = (M + 31) / 32, NZLs = NZL > 0 ? NZL : 1;
  const double* Dv = F64 + H->f_D;  const double* Dinv = F64 + H->f_Dinv;
  const double* Ev = F64 + H->f_E;  const double* Einv = F64 + H->f_Einv;
  const double c = H->c, cinv = H->cinv, sigma = H->sigma, alpha = st.alpha;
  const bool unscale = st.scaling && !st.scaled_termination;
  const double rho_in = H->rho, rho_eq = RHO_EQ_FACTOR * H->rho;
  const double rinv_in = 1.0 / rho_in, rinv_eq = 1.0 / rho_eq, rinv_loose = 1.0 / RHO_MIN;

  // ---- per-lane constants of the family (identical for every instance)
  // Pivot positions and the q/l/u accessor tables are read from the blob when needed (they would cost 30 registers).
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
  auto tab2 = [&](const int* T, int k) __attribute__((always_inline)) -> double2 {   // value of slots 0 and 1
    const int a = T[32 * k];
    if (a >= 0) { const double v = F64[a]; return make_double2(v, v); }
    return lds2(bv + 2 * (-a - 1));
  };
  auto q_of = [&](int k, int s) __attribute__((always_inline)) -> double { return sel(tab2(AQ, k), s); };
  auto l_of = [&](int k, int s) __attribute__((always_inline)) -> double { return sel(tab2(AL, k), s); };
  auto u_of = [&](int k, int s) __attribute__((always_inline)) -> double { return sel(tab2(AU, k), s); };
  auto rinv_of = [&](int k) __attribute__((always_inline)) -> double { return ((loosemask >> k) & 1u) ? rinv_loose : (((eqmask >> k) & 1u) ? rinv_eq : rinv_in); };
  auto rho_of = [&](int k) __attribute__((always_inline)) -> double { return ((loosemask >> k) & 1u) ? RHO_MIN : (((eqmask >> k) & 1u) ? rho_eq : rho_in); };

  // ---- per-slot state
  double x[2][NXL], z[2][NZLs], y[2][NZLs];
  double dx[2][NXL], dy[2][NZLs];
  int inst[2] = {-1, -1}, it[2] = {0, 0};
  bool active[2] = {false, false};
  bool exhausted = false;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
#pragma unroll
    for (int k = 0; k < NXL; ++k) { x[s][k] = 0.0; dx[s][k] = 0.0; }
#pragma unroll
    for (int k = 0; k < NZLs; ++k) { z[s][k] = 0.0; y[s][k] = 0.0; dy[s][k] = 0.0; }
  }
  // results of the last update_info, per slot
  double pri_res[2], dua_res[2], xPx[2], qx[2], nrm_z[2], nrm_Ax[2], nrm_q[2], nrm_Aty[2], nrm_Px[2];
  double s_rp[2], s_rd[2], s_z[2], s_Ax[2], s_q[2], s_Aty[2], s_Px[2];

  // ---- a1-a3: canonicalise the batched rows of slot s, scale them, detect constraint-type changes
  auto load_instance = [&](int s, int b) __attribute__((always_inline)) -> bool {
    const double* th = io.params + (size_t)b * H->npb;
    const int nbq = H->n_bq, nbc = H->n_bc;
    for (int r = lane; r < nbq; r += LANES) {
      const int i = U16[H->h_bq_row + r];
      double acc = F64[H->f_qbase + i];
      for (int e = I32[H->i_bq_ptr + r]; e < I32[H->i_bq_ptr + r + 1]; ++e)
        acc = fma(F64[H->f_bq_val + e], __ldg(th + U16[H->h_bq_col + e]), acc);
      bv[2 * r + s] = (Dv[i] * acc) * c;
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
      bv[2 * (nbq + r) + s] = al;
      bv[2 * (nbq + nbc + r) + s] = au;
      const bool loose = (al < -OSQP_INFTY * MIN_SCALING) && (au > OSQP_INFTY * MIN_SCALING);
      const bool eq = !loose && (au - al < RHO_TOL);
      mismatch |= ((loose ? 0 : (eq ? 2 : 1)) != (int)U16[H->h_ctype + j]);
    }
    __syncwarp();
    return __any_sync(FULL, mismatch);
  };

  // ---- residuals + norms of both slots' current iterates (update_info, auxil.c:564-629)
  auto update_info = [&]() __attribute__((always_inline)) {
#pragma unroll
    for (int k = 0; k < NXL; ++k) { const int i = lane + 32 * k; if (i < N) sts2(w2 + 2 * i, x[0][k], x[1][k]); }
#pragma unroll
    for (int k = 0; k < NZL; ++k) { const int j = lane + 32 * k; if (j < M) sts2(w2 + 2 * (N + j), y[0][k], y[1][k]); }
    __syncwarp();
    double m_rp[2] = {0, 0}, m_z[2] = {0, 0}, m_Ax[2] = {0, 0}, ms_rp[2] = {0, 0}, ms_z[2] = {0, 0}, ms_Ax[2] = {0, 0};
#pragma unroll
    for (int k = 0; k < NZL; ++k) {
      const int j = lane + 32 * k;
      if (j < M) {
        const double2 Ax2 = pair_ell_dot(I32 + H->i_ellA + 3 * k, F64, U16, w2, lane);
        const double e = unscale ? Einv[j] : 1.0;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const double Ax = sel(Ax2, s), rp = Ax - z[s][k];
          m_rp[s] = fmax(m_rp[s], fabs(e * rp)); m_z[s] = fmax(m_z[s], fabs(e * z[s][k])); m_Ax[s] = fmax(m_Ax[s], fabs(e * Ax));
          ms_rp[s] = fmax(ms_rp[s], fabs(rp)); ms_z[s] = fmax(ms_z[s], fabs(z[s][k])); ms_Ax[s] = fmax(ms_Ax[s], fabs(Ax));
        }
      }
    }
    double m_rd[2] = {0, 0}, m_q[2] = {0, 0}, m_Aty[2] = {0, 0}, m_Px[2] = {0, 0};
    double ms_rd[2] = {0, 0}, ms_q[2] = {0, 0}, ms_Aty[2] = {0, 0}, ms_Px[2] = {0, 0}, a_xPx[2] = {0, 0}, a_qx[2] = {0, 0};
#pragma unroll
    for (int k = 0; k < NXL; ++k) {
      const int i = lane + 32 * k;
      if (i < N) {
        const double2 Px2 = pair_ell_dot(I32 + H->i_ellP + 3 * k, F64, U16, w2, lane);
        const double2 Aty2 = (M > 0) ? pair_ell_dot(I32 + H->i_ellAt + 3 * k, F64, U16, w2, lane) : make_double2(0.0, 0.0);
        const double d = unscale ? Dinv[i] : 1.0;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const double Px = sel(Px2, s), Aty = sel(Aty2, s), qv = q_of(k, s);
          const double rd = qv + Px + Aty;
          m_rd[s] = fmax(m_rd[s], fabs(d * rd)); m_q[s] = fmax(m_q[s], fabs(d * qv));
          m_Aty[s] = fmax(m_Aty[s], fabs(d * Aty)); m_Px[s] = fmax(m_Px[s], fabs(d * Px));
          ms_rd[s] = fmax(ms_rd[s], fabs(rd)); ms_q[s] = fmax(ms_q[s], fabs(qv));
          ms_Aty[s] = fmax(ms_Aty[s], fabs(Aty)); ms_Px[s] = fmax(ms_Px[s], fabs(Px));
          a_xPx[s] = fma(x[s][k], Px, a_xPx[s]); a_qx[s] = fma(qv, x[s][k], a_qx[s]);
        }
      }
    }
    __syncwarp();
    const double cs = unscale ? cinv : 1.0;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      pri_res[s] = (M > 0) ? warp_max(m_rp[s]) : 0.0;
      nrm_z[s] = warp_max(m_z[s]); nrm_Ax[s] = warp_max(m_Ax[s]);
      dua_res[s] = cs * warp_max(m_rd[s]);
      nrm_q[s] = warp_max(m_q[s]); nrm_Aty[s] = warp_max(m_Aty[s]); nrm_Px[s] = warp_max(m_Px[s]);
      s_rp[s] = warp_max(ms_rp[s]); s_z[s] = warp_max(ms_z[s]); s_Ax[s] = warp_max(ms_Ax[s]);
      s_rd[s] = warp_max(ms_rd[s]); s_q[s] = warp_max(ms_q[s]); s_Aty[s] = warp_max(ms_Aty[s]); s_Px[s] = warp_max(ms_Px[s]);
      xPx[s] = warp_sum(a_xPx[s]); qx[s] = warp_sum(a_qx[s]);
    }
  };

  // ---- check_termination for slot s (auxil.c:681-786)
  auto check_termination = [&](int s, bool approximate) __attribute__((always_inline)) -> int {
    double ea = st.eps_abs, er = st.eps_rel, epi = st.eps_prim_inf, edi = st.eps_dual_inf;
    if (approximate) { ea *= 10; er *= 10; epi *= 10; edi *= 10; }
    if (pri_res[s] > OSQP_INFTY || dua_res[s] > OSQP_INFTY) return ST_NONCVX;
    const double cs = unscale ? cinv : 1.0;
    bool prim_ok, prim_inf = false, dual_inf = false;
    if (M == 0) prim_ok = true;
    else {
      prim_ok = pri_res[s] < ea + er * fmax(nrm_z[s], nrm_Ax[s]);
      if (!prim_ok) {               // is_primal_infeasible, auxil.c:361-424
        double dproj[NZLs];
        double nd = 0, lhs = 0;
#pragma unroll
        for (int k = 0; k < NZL; ++k) {
          const int j = lane + 32 * k;
          double d = 0.0;
          if (j < M) {
            d = dy[s][k];
            const double lv = l_of(k, s), uv = u_of(k, s);
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
            for (int k = 0; k < NZL; ++k) { const int j = lane + 32 * k; if (j < M) sts2(w2 + 2 * (N + j), dproj[k], dproj[k]); }
            __syncwarp();
            double mx = 0;
#pragma unroll
            for (int k = 0; k < NXL; ++k) {
              const int i = lane + 32 * k;
              if (i < N) {
                double v = pair_ell_dot(I32 + H->i_ellAt + 3 * k, F64, U16, w2, lane).x;
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
    const bool dual_ok = dua_res[s] < ea + er * cs * fmax(fmax(nrm_q[s], nrm_Aty[s]), nrm_Px[s]);
    if (!dual_ok) {                 // is_dual_infeasible, auxil.c:426-512
      double nd = 0, qd = 0;
#pragma unroll
      for (int k = 0; k < NXL; ++k) {
        const int i = lane + 32 * k;
        if (i < N) { nd = fmax(nd, fabs(unscale ? Dv[i] * dx[s][k] : dx[s][k])); qd = fma(q_of(k, s), dx[s][k], qd); }
      }
      nd = warp_max(nd);
      const double cost_scaling = unscale ? c : 1.0;
      if (nd > DIVISION_TOL) {
        qd = warp_sum(qd);
        if (qd < cost_scaling * edi * nd) {
#pragma unroll
          for (int k = 0; k < NXL; ++k) { const int i = lane + 32 * k; if (i < N) sts2(w2 + 2 * i, dx[s][k], dx[s][k]); }
          __syncwarp();
          double mx = 0;
#pragma unroll
          for (int k = 0; k < NXL; ++k) {
            const int i = lane + 32 * k;
            if (i < N) {
              double v = pair_ell_dot(I32 + H->i_ellP + 3 * k, F64, U16, w2, lane).x;
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
                double v = pair_ell_dot(I32 + H->i_ellA + 3 * k, F64, U16, w2, lane).x;
                if (unscale) v *= Einv[j];
                bad |= ((u_of(k, s) < OSQP_INFTY * MIN_SCALING) && (v > edi * nd)) ||
                       ((l_of(k, s) > -OSQP_INFTY * MIN_SCALING) && (v < -edi * nd));
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

  // ---- hand slot s to the tail kernel (own KKT factor needed)
  auto hand_off = [&](int s, double rho_new) __attribute__((always_inline)) {
    const int b = inst[s];
    int slot = -1;
    if (lane == 0) slot = atomicAdd(io.tail_count, 1);
    slot = __shfl_sync(FULL, slot, 0);
    if (lane == 0) { io.status[b] = ST_HANDOFF; io.iter[b] = it[s]; }
    if (slot < io.tail_capacity) {
      double* ts = io.tail_state + (size_t)slot * (N + 2 * M + 2);
#pragma unroll
      for (int k = 0; k < NXL; ++k) { const int i = lane + 32 * k; if (i < N) ts[i] = x[s][k]; }
#pragma unroll
      for (int k = 0; k < NZL; ++k) { const int j = lane + 32 * k; if (j < M) { ts[N + j] = z[s][k]; ts[N + M + j] = y[s][k]; } }
      if (lane == 0) { ts[N + 2 * M] = rho_new; ts[N + 2 * M + 1] = (double)it[s]; io.tail_ids[slot] = b; }
    }
  };

  // ---- store_solution / unscale / retrieval for slot s (a11, a12)
  auto finish = [&](int s, int status) __attribute__((always_inline)) {
    const int b = inst[s];
    const bool has_sol = !(status == ST_PINF || status == ST_PINF_INACC || status == ST_DINF ||
                           status == ST_DINF_INACC || status == ST_NONCVX);
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
    for (int k = 0; k < NXL; ++k) {
      const int i = lane + 32 * k;
      if (i < N) {
        const double xv = has_sol ? Dv[i] * x[s][k] : qnan;
        w2[2 * i] = xv;
        if (io.sol_x) io.sol_x[(size_t)b * N + i] = xv;
      }
    }
#pragma unroll
    for (int k = 0; k < NZL; ++k) {
      const int j = lane + 32 * k;
      if (j < M) {
        const double yv = has_sol ? (Ev[j] * y[s][k]) * cinv : qnan;
        w2[2 * (N + j)] = yv;
        if (io.sol_y) io.sol_y[(size_t)b * M + j] = yv;
      }
    }
    __syncwarp();
    if (io.prim) {
      const int np = H->n_prim;
      for (int k = lane; k < np; k += LANES) io.prim[(size_t)b * np + k] = w2[2 * U16[H->h_prim + k]];
    }
    if (io.dual) {
      const int nd = H->n_dual;
      for (int k = lane; k < nd; k += LANES) io.dual[(size_t)b * nd + k] = w2[2 * (N + U16[H->h_dual + k])];
    }
    __syncwarp();
    if (lane == 0) {
      double obj = (0.5 * xPx[s] + qx[s]);
      if (st.scaling) obj *= cinv;
      if (status == ST_PINF || status == ST_PINF_INACC) obj = OSQP_INFTY;
      else if (status == ST_DINF || status == ST_DINF_INACC) obj = -OSQP_INFTY;
      else if (status == ST_NONCVX) obj = qnan;
      else obj = (H->is_max ? -1.0 : 1.0) * (obj + H->d_const);
      io.obj_val[b] = obj; io.iter[b] = it[s]; io.status[b] = status;
      io.pri_res[b] = pri_res[s]; io.dua_res[b] = dua_res[s];
    }
  };

  for (;;) {
    // All warps of the CTA run the same ~50 KB of straight-line code per iteration; without this barrier they drift
    // apart and every warp misses the instruction cache on its own (ncu: stall_no_instruction 4.4 cycles/issue).
    // Iteration counts are multiples of check_termination for every instance, so the warps' check iterations coincide.
    if (!__syncthreads_or((int)(!exhausted || active[0] || active[1]))) break;
    // ---- refill empty slots from the global queue
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      while (!active[s] && !exhausted) {
        unsigned b = 0;
        if (lane == 0) b = atomicAdd(io.work_counter, 1u);
        b = __shfl_sync(FULL, b, 0);
        if (b >= (unsigned)io.B) { exhausted = true; break; }
        inst[s] = (int)b; it[s] = 0;
        const bool mismatch = load_instance(s, (int)b);
        // cold start (auxil.c:155-159) or warm start (osqp.c:929-953)
        if (st.warm_start && io.x0 != nullptr && io.y0 != nullptr) {
#pragma unroll
          for (int k = 0; k < NXL; ++k) {
            const int i = lane + 32 * k;
            if (i < N) { x[s][k] = Dinv[i] * io.x0[(size_t)b * N + i]; sts2(w2 + 2 * i, x[s][k], x[s][k]); }
          }
#pragma unroll
          for (int k = 0; k < NZL; ++k) { const int j = lane + 32 * k; if (j < M) y[s][k] = (Einv[j] * io.y0[(size_t)b * M + j]) * c; }
          __syncwarp();
#pragma unroll
          for (int k = 0; k < NZL; ++k) { const int j = lane + 32 * k; if (j < M) z[s][k] = pair_ell_dot(I32 + H->i_ellA + 3 * k, F64, U16, w2, lane).x; }
          __syncwarp();
        } else {
#pragma unroll
          for (int k = 0; k < NXL; ++k) x[s][k] = 0.0;
#pragma unroll
          for (int k = 0; k < NZLs; ++k) { z[s][k] = 0.0; y[s][k] = 0.0; }
        }
        if (mismatch) { hand_off(s, rho_in); continue; }     // a constraint changed type: needs its own factor
        if (st.max_iter <= 0) {                                 // degenerate setting: report the start point
          update_info();
          int status = check_termination(s, false);
          if (status == ST_UNSOLVED) { status = check_termination(s, true); if (status == ST_UNSOLVED) status = ST_MAXITER; }
          finish(s, status);
          continue;
        }
        active[s] = true;
      }
    }
    if (!active[0] && !active[1]) continue;    // nothing left for this warp: keep meeting the others at the barrier

    // which slots evaluate their residuals after this iteration
    bool chk[2], adp[2];
    bool any_chk = false;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int itn = it[s] + 1;
      const bool can_check = st.check_termination && (itn % st.check_termination == 0);
      adp[s] = active[s] && st.adaptive_rho && st.adaptive_rho_interval && (itn % st.adaptive_rho_interval == 0);
      chk[s] = active[s] && (can_check || itn == st.max_iter);
      any_chk |= chk[s] || adp[s];
    }

    // ---- one ADMM iteration for both slots (osqp.c:354-372)
#pragma unroll
    for (int k = 0; k < NXL; ++k) {
      const int i = lane + 32 * k;
      if (i < N) { const double2 qv = tab2(AQ, k); sts2(w2 + 2 * PX[32 * k], sigma * x[0][k] - qv.x, sigma * x[1][k] - qv.y); }
    }
#pragma unroll
    for (int k = 0; k < NZL; ++k) {
      const int j = lane + 32 * k;
      if (j < M) { const double ri = rinv_of(k); sts2(w2 + 2 * PZ[32 * k], z[0][k] - ri * y[0][k], z[1][k] - ri * y[1][k]); }
    }
    __syncwarp();
#ifdef CPG_FAM_GENERATED_SOLVE
    cpg_kkt_solve_gen(F64, I32, U16, w2, lane);
#else
    pair_kkt_solve<Fam::TRAIL>(H, I32, F64, U16, w2, lane);
#endif
#pragma unroll
    for (int k = 0; k < NXL; ++k) {
      const int i = lane + 32 * k;
      if (i < N) {
        const double2 xt = lds2(w2 + 2 * PX[32 * k]);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const double xn = alpha * sel(xt, s) + (1.0 - alpha) * x[s][k];
          dx[s][k] = xn - x[s][k];
          x[s][k] = xn;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < NZL; ++k) {
      const int j = lane + 32 * k;
      if (j < M) {
        const double ri = rinv_of(k), r = rho_of(k);
        const double2 nu = lds2(w2 + 2 * PZ[32 * k]);
        const double2 lv = tab2(AL, k), uv = tab2(AU, k);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const double zt = (z[s][k] - ri * y[s][k]) + ri * sel(nu, s);
          const double v = alpha * zt + (1.0 - alpha) * z[s][k];
          const double zn = fmin(fmax(v + ri * y[s][k], sel(lv, s)), sel(uv, s));
          const double d = r * (v - zn);
          dy[s][k] = d;
          y[s][k] += d;
          z[s][k] = zn;
        }
      }
    }
    __syncwarp();
    it[0] += active[0] ? 1 : 0;
    it[1] += active[1] ? 1 : 0;

    if (any_chk) {
      update_info();
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (!(chk[s] || adp[s])) continue;
        int status = ST_UNSOLVED;
        if (chk[s]) status = check_termination(s, false);
        if (status == ST_UNSOLVED && it[s] >= st.max_iter) {          // osqp.c:563-568
          status = check_termination(s, true);
          if (status == ST_UNSOLVED) status = ST_MAXITER;
        }
        if (status != ST_UNSOLVED) { finish(s, status); active[s] = false; continue; }
        if (adp[s]) {                                                 // adapt_rho decision (auxil.c:13-74)
          const double pn = s_rp[s] / (fmax(s_z[s], s_Ax[s]) + DIVISION_TOL);
          const double dn = s_rd[s] / (fmax(fmax(s_q[s], s_Aty[s]), s_Px[s]) + DIVISION_TOL);
          double r = rho_in * sqrt(pn / dn);
          r = fmin(fmax(r, RHO_MIN), RHO_MAX);
          if (r > rho_in * st.adaptive_rho_tolerance || r < rho_in / st.adaptive_rho_tolerance) {
            hand_off(s, r); active[s] = false;
          }
        }
      }
    }
  }
}

template <class Fam>
__global__ void __launch_bounds__(Fam::WARPS * 32, 1)
admm_pair_kernel(const uint8_t* __restrict__ blob_g, const BatchIO io, const Settings st) {
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
  double* w2 = reinterpret_cast<double*>(smem + Fam::BLOB_BYTES_PAD) + (size_t)warp * Fam::PAIR_STRIDE;
  double* bv = w2 + 2 * Fam::W_STRIDE;
  solve_pairs<Fam>(H, I32, F64, U16, w2, bv, lane, io, st);
}

}  // namespace cpgb200
