// grad_kernel.cuh -- batched backward pass through the QP (gradient=True), one instance per warp, sm_100a.
//
// Reference path restated:
//   cpg_osqp_gradient()       cvxpygen/templates/cpg_osqp_grad_compute.c.jinja2:432-531
//   K, K_true definition      cvxpygen/writer.py:361-369
//   un-canonicalisation       cvxpygen/writer.py:268-303     (dp = sum_id map_id' d(id); vector ids here)
//   cpg_update_d<var>         cvxpygen/writer.py:222-230     (scatter of the user-variable gradients into dx)
// Per instance, with (x, y) the canonical solution of the forward pass and dx the upstream gradient:
//   active set  a_j = sign(y_j) if |y_j| > 1e-12 else 0
//   K r = [dx; 0],  K = [[P + 1e-6 I, A_act'], [A_act, -1e-6 I]], inactive rows replaced by the pivot -1
//   3 x iterative refinement against K_true = [[P, A_act'], [A_act, 0]]
//   dq = -r_x,  dl_j = r_{n+j} [a_j = -1],  du_j = r_{n+j} [a_j = +1],  dtheta = Mq' dq + Ml' dl + Mu' du
// The reference keeps ONE factor alive and moves it from one call's active set to the next with rank-1
// up/down-dates (:157-324) -- an optimisation for sequential solves with no analogue in an independent batch.
// Here every instance factors its own K numerically on the family's symbolic pattern with the same level-wise
// right-looking scheme as the tail kernel (tail_factor / tail_solve in admm_kernel.cuh, tables of
// offline/refactor.py), which gives the same r up to rounding (3 refinement steps on both sides).
// One stateless difference: the reference's dl/du split follows a[] carried over from EARLIER calls (a row that
// flips from upper- to lower-active without passing through inactive keeps its old label, :437-449); here the
// split follows sign(y) of the instance itself.  dl + du and dtheta are unaffected for l = u rows.
// Families with per-instance MATRIX parameters (rows a16 + f2; template flag MATPAR): the instance's unscaled P and A are
// canonicalised from its parameter row first (what cpg_P_to_K / cpg_A_to_K + cpg_ldl_numeric do when P / A are outdated,
// cvxpygen/writer.py:240-263), K is assembled through the slot maps of the matrix blob, the refinement products run over
// the index tables, and the matrix gradients  dP_k = -1/2 (r_i x_j + x_i r_j),  dA_k = -(r_{n+i} x_j + y_i r_j) [row active]
// (:513-529) are folded into dtheta through the transposed maps of the P / A entries (writer.py:292-303).
#pragma once
#include "admm_kernel.cuh"
#if CPG_FAM_MATPAR
#include "matpar_kernel.cuh"
#endif

namespace cpgb200 {

constexpr double GRAD_ACTIVE_TOL = 1e-12;

struct GradIO {
  const double* sol_y;     // (B, m) canonical dual of the forward pass (unscaled)
  const double* dprim;     // (B, n_prim) upstream gradient w.r.t. the user-level primal variables (prim layout)
  double* dparams;         // (B, npb) gradient w.r.t. the batched user parameters
  double* dq;              // optional (B, n) / (B, m) canonical gradients
  double* dl;
  double* du;
  const double* S0;        // (n_slots) regularised KKT in slot order (global memory)
  int B;
  // matrix-parameter families only
  const double* params;    // (B, npb) the instances' parameter rows (their P and A are canonicalised from them)
  const double* sol_x;     // (B, n) canonical primal solution (enters dP and dA)
  double* dP;              // optional (B, nnzP) / (B, nnzA) canonical matrix gradients
  double* dA;
  const uint8_t* mblob;    // matrix tables (global)
  double* a_scratch;       // per-warp slices for the entries of A
};

template <class Fam, bool MATPAR = false>
__global__ void __launch_bounds__(Fam::GRAD_WARPS * 32, 1)
qp_grad_kernel(const uint8_t* __restrict__ gblob_g, const uint8_t* __restrict__ tail_blob_g, const GradIO io) {
  constexpr int N = Fam::N, M = Fam::M, NK = N + M, NXL = (N + 31) / 32, NZL = (M + 31) / 32, NZLs = NZL > 0 ? NZL : 1;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t total = reinterpret_cast<const CpgGradHeader*>(gblob_g)->total_bytes;
  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (tid == 0) {                      // stage the gradient constants with TMA bulk copies
    mbar_expect_tx(&bar, total);
    constexpr uint32_t CHUNK = 32768;
    for (uint32_t off = 0; off < total; off += CHUNK)
      tma_bulk_g2s(smem + off, gblob_g + off, (total - off < CHUNK) ? (total - off) : CHUNK, &bar);
  }
  mbar_wait(&bar, 0);
  const CpgGradHeader* H = reinterpret_cast<const CpgGradHeader*>(smem);
  const int* I32 = reinterpret_cast<const int*>(smem + H->off_i32);
  const double* F64 = reinterpret_cast<const double*>(smem + H->off_f64);
  const uint16_t* U16 = reinterpret_cast<const uint16_t*>(smem + H->off_u16);
  const TailView tv = make_tail_view(tail_blob_g);
  // per-warp shared memory: S | w (pivot order) | r (natural order) | rhs_x (n) | y (m)
  double* S = reinterpret_cast<double*>(smem + Fam::GBLOB_BYTES_PAD) + (size_t)warp * Fam::GRAD_STRIDE;
  double* w = S + Fam::S_STRIDE;
  double* r = w + Fam::W_STRIDE;
  double* rhsx = r + Fam::W_STRIDE;
  double* ys = rhsx + ((N + 1) & ~1);
  const int n_slots = H->n_slots;
#if CPG_FAM_MATPAR
  double* xs = ys + ((M + 1) & ~1);          // canonical x of the instance
  MatCtx mc;
  if constexpr (MATPAR) {
    mc.mv = make_mat_view(io.mblob);
    mc.Pv = xs + ((N + 1) & ~1);
    mc.Av = io.a_scratch + ((size_t)blockIdx.x * Fam::GRAD_WARPS + warp) * Fam::MAT_G_STRIDE;
  }
#endif
  auto dotA = [&](int k) -> double {
#if CPG_FAM_MATPAR
    if constexpr (MATPAR) return ellx_dot(mc.mv.I32 + mc.mv.H->i_ixA + 3 * k, mc.mv.U16, mc.Av, r, lane);
#endif
    return ell_dot(I32 + H->i_ellA + 3 * k, F64, U16, r, lane);
  };
  auto dotAt = [&](int k) -> double {
#if CPG_FAM_MATPAR
    if constexpr (MATPAR) return ellx_dot(mc.mv.I32 + mc.mv.H->i_ixAt + 3 * k, mc.mv.U16, mc.Av, r, lane);
#endif
    return ell_dot(I32 + H->i_ellAt + 3 * k, F64, U16, r, lane);
  };
  auto dotP = [&](int k) -> double {
#if CPG_FAM_MATPAR
    if constexpr (MATPAR) return ellx_dot(mc.mv.I32 + mc.mv.H->i_ixP + 3 * k, mc.mv.U16, mc.Pv, r, lane);
#endif
    return ell_dot(I32 + H->i_ellP + 3 * k, F64, U16, r, lane);
  };

  for (int b = blockIdx.x * Fam::GRAD_WARPS + warp; b < io.B; b += gridDim.x * Fam::GRAD_WARPS) {
    // ---- inputs: dual solution -> active set; upstream gradient scattered into canonical dx (cpg_update_d<var>)
    unsigned act = 0u;
#pragma unroll
    for (int k = 0; k < NZL; ++k) {
      const int j = lane + 32 * k;
      if (j < M) {
        const double yv = io.sol_y[(size_t)b * M + j];
        ys[j] = yv;
        if (yv < -GRAD_ACTIVE_TOL || yv > GRAD_ACTIVE_TOL) act |= 1u << k;
      }
    }
    for (int i = lane; i < N; i += LANES) rhsx[i] = 0.0;
#if CPG_FAM_MATPAR
    if constexpr (MATPAR) {
      // this instance's K_reg = [[P + reg I, A'], [A, -reg I]] on the symbolic pattern (cvxpygen/writer.py:361-364)
      matpar_canon(mc, io.params + (size_t)b * H->npb, lane);
      for (int i = lane; i < N; i += LANES) xs[i] = io.sol_x[(size_t)b * N + i];
      for (int i = lane; i < n_slots; i += LANES) S[i] = 0.0;
      __syncwarp();
      for (int i = lane; i < N; i += LANES) S[U16[H->h_pinvx + i]] = 1e-6;
      for (int j = lane; j < M; j += LANES) S[U16[H->h_pinvz + j]] = -1e-6;
      __syncwarp();
      const CpgMatHeader* MH = mc.mv.H;
      for (int e = lane; e < MH->nnzP; e += LANES) S[__ldg(mc.mv.U16 + MH->h_Pslot + e)] += mc.Pv[e];
      for (int e = lane; e < MH->nnzA; e += LANES) S[__ldg(mc.mv.U16 + MH->h_Aslot + e)] = mc.Av[e];
    } else
#endif
    {
      for (int i = lane; i < n_slots; i += LANES) S[i] = __ldg(io.S0 + i);
    }
    __syncwarp();
    {
      const int np = H->n_prim;
      const double* dp = io.dprim + (size_t)b * np;
      for (int k = lane; k < np; k += LANES) rhsx[U16[H->h_prim + k]] = dp[k];   // user variables do not overlap
    }
    // ---- K of this instance: inactive rows lose their A entries and get the pivot -1 (cpg_ldl_delete semantics)
#pragma unroll
    for (int k = 0; k < NZL; ++k) {
      const int j = lane + 32 * k;
      if (j < M && !((act >> k) & 1u)) {
        for (int e = I32[H->i_arow_ptr + j]; e < I32[H->i_arow_ptr + j + 1]; ++e) S[U16[H->h_arow_slot + e]] = 0.0;
        S[U16[H->h_pinvz + j]] = -1.0;
      }
    }
    __syncwarp();
    tail_factor(tv, S, lane);
    // ---- first solve: r = K^-1 [dx; 0]
#pragma unroll
    for (int k = 0; k < NXL; ++k) { const int i = lane + 32 * k; if (i < N) w[U16[H->h_pinvx + i]] = rhsx[i]; }
#pragma unroll
    for (int k = 0; k < NZL; ++k) { const int j = lane + 32 * k; if (j < M) w[U16[H->h_pinvz + j]] = 0.0; }
    __syncwarp();
    tail_solve(tv, S, w, lane);
#pragma unroll
    for (int k = 0; k < NXL; ++k) { const int i = lane + 32 * k; if (i < N) r[i] = w[U16[H->h_pinvx + i]]; }
#pragma unroll
    for (int k = 0; k < NZL; ++k) { const int j = lane + 32 * k; if (j < M) r[N + j] = ((act >> k) & 1u) ? w[U16[H->h_pinvz + j]] : 0.0; }
    __syncwarp();
    // ---- three steps of iterative refinement against the exact K_true (:456-490)
    for (int itr = 0; itr < 3; ++itr) {
      double dxk[NXL], dzk[NZLs];
#pragma unroll
      for (int k = 0; k < NXL; ++k) {
        const int i = lane + 32 * k;
        dxk[k] = 0.0;
        if (i < N) dxk[k] = rhsx[i] - dotP(k) - ((M > 0) ? dotAt(k) : 0.0);
      }
#pragma unroll
      for (int k = 0; k < NZL; ++k) {
        const int j = lane + 32 * k;
        dzk[k] = 0.0;
        if (j < M && ((act >> k) & 1u)) dzk[k] = -dotA(k);
      }
#pragma unroll
      for (int k = 0; k < NXL; ++k) { const int i = lane + 32 * k; if (i < N) w[U16[H->h_pinvx + i]] = dxk[k]; }
#pragma unroll
      for (int k = 0; k < NZL; ++k) { const int j = lane + 32 * k; if (j < M) w[U16[H->h_pinvz + j]] = dzk[k]; }
      __syncwarp();
      tail_solve(tv, S, w, lane);
#pragma unroll
      for (int k = 0; k < NXL; ++k) { const int i = lane + 32 * k; if (i < N) r[i] += w[U16[H->h_pinvx + i]]; }
#pragma unroll
      for (int k = 0; k < NZL; ++k) { const int j = lane + 32 * k; if (j < M && ((act >> k) & 1u)) r[N + j] += w[U16[H->h_pinvz + j]]; }
      __syncwarp();
    }
    // ---- canonical gradients and un-canonicalisation
    if (io.dq) for (int i = lane; i < N; i += LANES) io.dq[(size_t)b * N + i] = -r[i];
    if (io.dl || io.du) {
      for (int j = lane; j < M; j += LANES) {
        const double yv = ys[j], rv = r[N + j];
        if (io.dl) io.dl[(size_t)b * M + j] = (yv < -GRAD_ACTIVE_TOL) ? rv : 0.0;
        if (io.du) io.du[(size_t)b * M + j] = (yv > GRAD_ACTIVE_TOL) ? rv : 0.0;
      }
    }
#if CPG_FAM_MATPAR
    if constexpr (MATPAR) {
      const CpgMatHeader* MH = mc.mv.H;
      if (io.dP) for (int e = lane; e < MH->nnzP; e += LANES) {
        const int i = __ldg(mc.mv.U16 + MH->h_Prow + e), j = __ldg(mc.mv.U16 + MH->h_Pcol + e);
        io.dP[(size_t)b * MH->nnzP + e] = -0.5 * (r[i] * xs[j] + xs[i] * r[j]);
      }
      if (io.dA) for (int e = lane; e < MH->nnzA; e += LANES) {
        const int i = __ldg(mc.mv.U16 + MH->h_Arow + e), j = __ldg(mc.mv.U16 + MH->h_Acol + e);
        const double yv = ys[i];
        io.dA[(size_t)b * MH->nnzA + e] = (yv < -GRAD_ACTIVE_TOL || yv > GRAD_ACTIVE_TOL) ? -(r[N + i] * xs[j] + yv * r[j]) : 0.0;
      }
    }
#endif
    if (io.dparams) {
      const int npb = H->npb;
      for (int cidx = lane; cidx < npb; cidx += LANES) {
        double acc = 0.0;
        for (int e = I32[H->i_tptr + cidx]; e < I32[H->i_tptr + cidx + 1]; ++e) {
          const int kind = U16[H->h_tkind + e], idx = U16[H->h_tidx + e];
          const double v = F64[H->f_tval + e];
          if (kind == 0) acc -= v * r[idx];                                           // dq = -r_x
          else if (kind == 1) { if (ys[idx] < -GRAD_ACTIVE_TOL) acc += v * r[N + idx]; }   // dl
          else if (kind == 2) { if (ys[idx] > GRAD_ACTIVE_TOL) acc += v * r[N + idx]; }    // du
#if CPG_FAM_MATPAR
          else if (MATPAR && kind == 3) {                                                  // dP entry idx
            const int i = __ldg(mc.mv.U16 + mc.mv.H->h_Prow + idx), j = __ldg(mc.mv.U16 + mc.mv.H->h_Pcol + idx);
            acc += v * (-0.5 * (r[i] * xs[j] + xs[i] * r[j]));
          } else if (MATPAR && kind == 4) {                                                // dA entry idx (active rows only)
            const int i = __ldg(mc.mv.U16 + mc.mv.H->h_Arow + idx), j = __ldg(mc.mv.U16 + mc.mv.H->h_Acol + idx);
            const double yv = ys[i];
            if (yv < -GRAD_ACTIVE_TOL || yv > GRAD_ACTIVE_TOL) acc -= v * (r[N + i] * xs[j] + yv * r[j]);
          }
#endif
        }
        io.dparams[(size_t)b * npb + cidx] = acc;
      }
    }
    __syncwarp();
  }
}

}  // namespace cpgb200
