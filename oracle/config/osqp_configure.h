/* oracle/config/osqp_configure.h -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * Hand-written stand-in for the header that OSQP 0.6.2's cmake step would
 * generate from configure/osqp_configure.h.in (reference:
 * cvxpygen/solvers/osqp-python/osqp_sources/configure/osqp_configure.h.in:1-49).
 * Choices mirror what cvxpygen's generated code uses: double floats, `int`
 * indices (cvxpygen/solvers/osqp.py:95), no printing, and NO profiling timer so
 * that adaptive_rho_interval is the deterministic 4*check_termination = 100
 * (osqp_sources/src/osqp.c:267-279).
 */
#ifndef OSQP_CONFIGURE_H
#define OSQP_CONFIGURE_H
#define IS_LINUX
/* not defined on purpose: DEBUG, EMBEDDED, PRINTING, PROFILING, CTRLC, DFLOAT, DLONG, ENABLE_MKL_PARDISO */
#endif
