// matpar_kernel.cuh -- per-instance MATRIX parameters, one instance per warp, sm_100a.
//
// When a batched user parameter enters the canonical P or A, the reference's generated solve goes through
//   cpg_canonicalize_P / _A            cvxpygen/utils.py:279-294            (P->x, A->x = map * params)
//   osqp_update_data_mat               cvxpygen/solvers/osqp.py:20-33        (0.6.2: osqp_update_P_A, osqp.c:1158-1264)
//     unscale_data                     scaling.c:160-175
//     scale_data                       scaling.c:44-156   (10 Ruiz iterations + cost normalisation, from D = E = c = 1)
//     update_linsys_solver_matrices    qdldl_interface.c:378-394  (update_KKT_P / update_KKT_A, kkt.c:184-212; QDLDL_factor)
//   osqp_update_data_vec               (q <- c D q; l,u <- E l, E u; update_rho_vec, auxil.c:100-142)
//   osqp_solve
// for every instance.  Batch semantics as everywhere in this backend: each instance starts from the pristine post-setup
// workspace, so the linear cost scale_data sees is the generation-time q after its unscale round trip (blob: f_q_un).
//
// Here one warp owns one instance end to end:
//   1. canonicalise its P and A entries (base + map * theta, lane-interleaved ELL over the ENTRIES),
//   2. equilibrate them in shared memory exactly in the reference's operation order (column / row infinity norms through
//      index tables, limit_scaling, sqrt, reciprocal, pre/post-multiplication, left-to-right mean of the column norms of
//      P) -- D, E and c come out bit-identical to scale_data's,
//   3. assemble K(rho_vec) on the family's symbolic pattern and factor it numerically (tail_factor),
//   4. run the ADMM loop of solve_instance<Fam, 2>: triangular solves over this instance's factor, residual products
//      through the index tables, in-place re-factorisation when rho adapts.
// HBM traffic per instance: its parameter row in, its solution rows out; the tables are shared by all warps (L2).
#pragma once
#include "admm_kernel.cuh"

namespace cpgb200 {

__device__ __forceinline__ double limit_scaling1(double v) {      // scaling.c:7-14
  v = v < MIN_SCALING ? 1.0 : v;
  return v > 1e4 ? 1e4 : v;
}

// step 1: cpg_canonicalize_P / _A for this instance (UNSCALED entries into mc.Pv / mc.Av; slot nnz = 0 for padding)
__device__ __forceinline__ void matpar_canon(MatCtx& mc, const double* __restrict__ th, const int lane) {
  const CpgMatHeader* H = mc.mv.H;
  const int* I32 = mc.mv.I32; const double* F64 = mc.mv.F64; const uint16_t* U16 = mc.mv.U16;
  auto canon = [&](double* out, int nnz, int i_ell, int f_base) {
    for (int e0 = 0, blk = 0; e0 < nnz; e0 += LANES, ++blk) {
      const int e = e0 + lane;
      const int* tab = I32 + i_ell + 3 * blk;
      const int K = __ldg(tab), fo = __ldg(tab + 1), uo = __ldg(tab + 2);
      if (e < nnz) {
        double acc = __ldg(F64 + f_base + e);
        for (int kk = 0; kk < K; ++kk)
          acc = fma(__ldg(F64 + fo + kk * LANES + lane), __ldg(th + __ldg(U16 + uo + kk * LANES + lane)), acc);
        out[e] = acc;
      }
    }
    if (lane == 0) out[nnz] = 0.0;
  };
  canon(mc.Pv, H->nnzP, H->i_ellMP, H->f_Pbase);
  canon(mc.Av, H->nnzA, H->i_ellMA, H->f_Abase);
  __syncwarp();
}

// steps 1-2; on exit mc.Av / mc.Pv hold the scaled matrices, mc.D/Dinv/E/Einv/c/cinv the scalings.  w: >= n+m doubles scratch.
template <class Fam>
__device__ void matpar_prepare(MatCtx& mc, const double* __restrict__ th, double* __restrict__ w, const int lane,
                               const int scaling, double* __restrict__ S) {
  constexpr int N = Fam::N, M = Fam::M, NXL = (N + 31) / 32, NZL = (M + 31) / 32, NZLs = NZL > 0 ? NZL : 1;
  const CpgMatHeader* H = mc.mv.H;
  const int* I32 = mc.mv.I32; const double* F64 = mc.mv.F64; const uint16_t* U16 = mc.mv.U16;
  const int nnzP = H->nnzP, nnzA = H->nnzA;
  // The ten Ruiz passes read every entry of A twice and rewrite it once per pass.  Their home is the warp's slice of the global
  // scratch (an L2 round trip per access, behind an index that is itself an L2 load: 13 % of the kernel's stall samples for 5 %
  // of its instructions, profiles/r2_matpar_v8_ncu_summary.md); the factor storage S is idle until the assembly, so the entries
  // are equilibrated THERE and copied to their home once at the end.
  double* const Ahome = mc.Av;
  constexpr bool STAGE_A = Fam::MAT_A_STRIDE <= Fam::S_STRIDE;
  if (STAGE_A) mc.Av = S;
  matpar_canon(mc, th, lane);
  // ---- 2. scale_data
  double d[NXL], e_[NZLs], q[NXL];
#pragma unroll
  for (int k = 0; k < NXL; ++k) { const int i = lane + 32 * k; d[k] = 1.0; q[k] = (i < N) ? __ldg(F64 + H->f_q_un + i) : 0.0; }
#pragma unroll
  for (int k = 0; k < NZLs; ++k) e_[k] = 1.0;
  double c = 1.0;
  for (int itr = 0; itr < scaling; ++itr) {
    // compute_inf_norm_cols_KKT (scaling.c:28-42) -> 1/sqrt(limit_scaling(.))
#pragma unroll
    for (int k = 0; k < NXL; ++k) {
      const int i = lane + 32 * k;
      if (i < N) {
        double nr = ellx_absmax(I32 + H->i_ixP + 3 * k, U16, mc.Pv, lane);
        if (M > 0) nr = fmax(nr, ellx_absmax(I32 + H->i_ixAt + 3 * k, U16, mc.Av, lane));
        const double dt = 1.0 / sqrt(limit_scaling1(nr));
        w[i] = dt; d[k] *= dt; q[k] *= dt;
      }
    }
#pragma unroll
    for (int k = 0; k < NZL; ++k) {
      const int j = lane + 32 * k;
      if (j < M) {
        const double et = 1.0 / sqrt(limit_scaling1(ellx_absmax(I32 + H->i_ixA + 3 * k, U16, mc.Av, lane)));
        w[N + j] = et; e_[k] *= et;
      }
    }
    __syncwarp();
    // P <- Dt P Dt, A <- Et A Dt (mat_premult_diag then mat_postmult_diag)
#if CPG_EQ_UNROLL
#pragma unroll 2
#endif
    for (int e = lane; e < nnzP; e += LANES)
      mc.Pv[e] = (mc.Pv[e] * w[__ldg(U16 + H->h_Prow + e)]) * w[__ldg(U16 + H->h_Pcol + e)];
#if CPG_EQ_UNROLL
#pragma unroll 4                 // (entries and indices come from L2: several elements in flight per lane)
#endif
    for (int e = lane; e < nnzA; e += LANES)
      mc.Av[e] = (mc.Av[e] * w[N + __ldg(U16 + H->h_Arow + e)]) * w[__ldg(U16 + H->h_Acol + e)];
    __syncwarp();
    // cost normalisation (scaling.c:112-142): mean of the column norms of P (summed left to right like vec_mean), |q|_inf
    double mq = 0.0;
#pragma unroll
    for (int k = 0; k < NXL; ++k) {
      const int i = lane + 32 * k;
      if (i < N) { w[i] = ellx_absmax(I32 + H->i_ixP + 3 * k, U16, mc.Pv, lane); mq = fmax(mq, fabs(q[k])); }
    }
    __syncwarp();
    double mean = 0.0;
    for (int i = 0; i < N; ++i) mean += w[i];
    mean /= (double)N;
    const double nq = limit_scaling1(warp_max(mq));
    const double ct = 1.0 / limit_scaling1(fmax(mean, nq));
    __syncwarp();
    for (int e = lane; e < nnzP; e += LANES) mc.Pv[e] *= ct;
#pragma unroll
    for (int k = 0; k < NXL; ++k) q[k] *= ct;
    c *= ct;
    __syncwarp();
  }
#pragma unroll
  for (int k = 0; k < NXL; ++k) { const int i = lane + 32 * k; if (i < N) { mc.D[i] = d[k]; mc.Dinv[i] = 1.0 / d[k]; } }
#pragma unroll
  for (int k = 0; k < NZL; ++k) { const int j = lane + 32 * k; if (j < M) { mc.E[j] = e_[k]; mc.Einv[j] = 1.0 / e_[k]; } }
  mc.c = c; mc.cinv = 1.0 / c;
  if (STAGE_A) {
    for (int e = lane; e <= nnzA; e += LANES) Ahome[e] = S[e];
    mc.Av = Ahome;
  }
  __syncwarp();
}

// One persistent CTA per SM; every warp pulls instance numbers from the global counter.  Nothing is staged: the family's
// compact blob, the refactorisation tables and the matrix tables are read from global memory (L2-resident, shared by all
// warps) so that shared memory is left to the per-instance state that every ADMM iteration touches: w | S | Pv.  The scaled
// entries of A and the scalings D, 1/D, E, 1/E -- read by the equilibration, the assembly, the residual checks every 25
// iterations and the epilogue only -- live in a per-warp slice of a global scratch buffer (a few MB in total:
// L2-resident), which is what lets five instead of three warps share an SM at MPC-12/4/10 with dense dynamics.
template <class Fam>
__global__ void __launch_bounds__(Fam::MAT_WARPS * 32, 1)
admm_matpar_kernel(const uint8_t* __restrict__ cblob_g, const uint8_t* __restrict__ tail_blob_g,
                   const uint8_t* __restrict__ mblob_g, double* __restrict__ a_scratch, const BatchIO io, const Settings st) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const CpgBlobHeader* H = reinterpret_cast<const CpgBlobHeader*>(cblob_g);
  const int* I32 = reinterpret_cast<const int*>(cblob_g + H->off_i32);
  const double* F64 = reinterpret_cast<const double*>(cblob_g + H->off_f64);
  const uint16_t* U16 = reinterpret_cast<const uint16_t*>(cblob_g + H->off_u16);
  double* wbase = reinterpret_cast<double*>(smem) + (size_t)warp * Fam::MAT_STRIDE;
  MatCtx mc;
  mc.mv = make_mat_view(mblob_g);
  TailArgs ta;
  ta.tv = make_tail_view(tail_blob_g);
  ta.S = wbase + Fam::W_STRIDE;
  ta.state = nullptr;
  ta.mc = &mc;
  mc.Av = a_scratch + ((size_t)blockIdx.x * Fam::MAT_WARPS + warp) * Fam::MAT_G_STRIDE;
  mc.Pv = ta.S + Fam::S_STRIDE;
  constexpr int NP = (Fam::N + 1) & ~1, MP = (Fam::M + 1) & ~1;
  mc.D = mc.Av + Fam::MAT_A_STRIDE;        // the scalings are read by the prologue, the checks and the epilogue only
  mc.Dinv = mc.D + NP; mc.E = mc.Dinv + NP; mc.Einv = mc.E + MP;
  for (;;) {
    int b = 0;
    if (lane == 0) b = (int)atomicAdd(io.work_counter, 1u);
    b = __shfl_sync(FULL, b, 0);
    if (b >= io.B) break;
    matpar_prepare<Fam>(mc, io.params + (size_t)b * H->npb, wbase, lane, st.scaling, ta.S);
    solve_instance<Fam, 2>(H, I32, F64, U16, wbase, lane, b, io, st, &ta);
    __syncwarp();
  }
}

}  // namespace cpgb200
