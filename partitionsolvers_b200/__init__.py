"""partitionsolvers_b200 — the consecutive-partitions DP of pehlivanian/PartitionSolvers on one B200 (sm_100a).

Layout: ``csrc/`` CUDA kernels and the C ABI (``include/ps_b200.h``), ``cxx/`` the C++ facade and the ``proto``
module, ``capi.py`` the ctypes binding used by tests and bench.py, ``host.py`` / ``solverSWIG.py`` /
``randomization.py`` the mirrors of the reference's Python callers, ``multigpu.py`` the row-sharded driver.
Nothing here falls back to a CPU path: every entry point raises when ``_lib/libps_b200.so`` is missing
(build it with ``python __graft_entry__.py``).  Importing the package itself loads nothing.
"""
