// tests/emu/ipm_emu.cpp -- TEST INFRASTRUCTURE.  Compiles the IPM-CUDA kernel source as plain host C++
// (-DCPG_IPM_HOST_EMU: every barrier-separated phase runs its threads one after the other) so that the phase logic and
// the generated tables can be checked against the oracle on a machine without a GPU.  Never linked into the product.
#define CPG_IPM_HOST_EMU 1
#include "cpg_ipm_family.h"
#include "ipm_kernel.cuh"

extern "C" int ipm_emu_smem_bytes() { return int(cpgipm::SMEM_BYTES); }
extern "C" long ipm_emu_count(int i) { return cpgipm::g_count[i]; }   // 0 ldl solves, 1 factorisations, 2 barriers, 3 KKT solves

extern "C" int ipm_emu_solve(const unsigned char* sblob, const unsigned char* gblob, int B, const double* params,
                             double* prim, double* dual, double* x, double* y, double* z, double* s,
                             double* obj, int* iter, int* status, double* pres, double* dres, int maxit) {
  using namespace cpgipm;
  std::vector<double> raw(SMEM_BYTES / 8 + 2, 0.0);
  unsigned char* base = reinterpret_cast<unsigned char*>(raw.data());
  std::memcpy(reinterpret_cast<double*>(base) + O_AG, sblob, size_t(NNZM + 1) * 8);
  std::memcpy(reinterpret_cast<double*>(base) + O_F64_END, sblob + IPM_SB_U32_OFF, size_t(U32_COUNT) * 4 + size_t(U16_COUNT) * 2);
  std::vector<double> best(BEST_STRIDE);
  Solver sv;
  sv.sm.base_ = base; sv.gm = make_gm(gblob); sv.rb = 0;
  sv.stg = IpmSettings{maxit, 0, 1e-8, 1e-8, 1e-8, 1e-4, 5e-5, 5e-5};
  IpmIO io{B, params, prim, dual, x, y, z, s, obj, iter, status, pres, dres, best.data(), nullptr};
  for (int i = 0; i < B; ++i) sv.solve_instance(i, io, best.data());
  return 0;
}
