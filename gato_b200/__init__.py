"""gato_b200 — B200-native (sm_100a) batched SQP trajectory-optimisation solver, a drop-in for the BSQP solve path of
A2R-Lab/GATO.  Native code: gato_b200/csrc (CUDA kernels + C ABI, include/gato_b200.h); Python: `native` (ctypes binding),
`bsqp` (mirror of the reference's Python modules), `workloads` (synthetic inputs), `sharding` (multi-GPU plumbing)."""
__version__ = "0.1.0"
