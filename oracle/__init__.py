"""oracle/ -- TEST INFRASTRUCTURE, not product code.

CPU checkers for the batched-solve hot path:
  * ref_osqp.py   ctypes binding of oracle/_ref/libosqp_ref.so (the UNMODIFIED
                  vendored OSQP 0.6.2 compiled from /root/reference, see Makefile)
  * admm_numpy.py numpy restatement of OSQP's ADMM (cites reference file:line)
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package.  The product package
(cvxpygen_b200) never does.
"""
