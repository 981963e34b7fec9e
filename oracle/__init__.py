"""TEST INFRASTRUCTURE.  CPU restatement of the reference's lattice algorithm (the oracle) and the
recipe that compiles the reference's own device kernels (oracle/_ref).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import this package;
the product (lattice_net_b200/) never does."""
