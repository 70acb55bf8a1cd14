"""direct_b200 -- B200-native batched IPDDP trajectory optimiser behind ntu-caokun/DIRECT's
ddpTrajOptimizer API.  The product is libdirect_ddp_b200.so (include/direct_ddp.h); this package
holds its CUDA sources (csrc/), the replacement C++ translation unit for the reference node (host/),
the ctypes binding used by tests and bench.py (capi.py) and the synthetic workload generator."""
from .problems import STAGE0, STAGE1, TIME_POWER, ProblemBatch, make_batch  # noqa: F401
