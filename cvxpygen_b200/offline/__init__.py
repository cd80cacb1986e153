"""Generation-time (host, numpy/scipy) half of the ADMM-CUDA backend: everything OSQP does in
osqp_setup -- equilibration, rho vector, KKT assembly, ordering, LDL' -- plus the
warp-level solve schedule the sm_100a kernel executes.  Runs once per problem family."""
