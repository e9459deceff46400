"""madnlp.jl_b200 -- B200-native KKT hot path for MadNLP-style interior-point solvers.

Only what the hot path needs lives here:
  csrc/           hand-written sm_100a CUDA kernels + the C ABI (include/b200kkt.h) -> libb200kkt.so
  capi.py         ctypes binding of that ABI (the same boundary Julia would `ccall`)
  linear_solvers  mirror of MadNLP's AbstractLinearSolver surface (B200SparseSolver, B200DenseSolver)
  kkt             mirror of the AbstractKKTSystem surface (SparseKKTSystem, SparseCondensedKKTSystem,
                  DenseCondensedKKTSystem, UnreducedKKTVector)
  richardson, ipm the refinement loop and the `regular!` call-order replay used for the IPM-level metric
  workloads       synthetic generators for the configurations named in BASELINE.json
  julia/          the Julia shim a MadNLP.jl maintainer would add (cannot be run in this image)

Importing this package requires the built shared library; there is no CPU fallback.
"""
from . import capi  # noqa: F401  (fails loudly if libb200kkt.so is missing)
from . import workloads  # noqa: F401


def __getattr__(name):
    # torch-dependent modules are imported lazily so that CPU-only tooling (ABI checks, symbolic analysis)
    # does not pay for `import torch`.
    import importlib
    if name in ("kkt", "linear_solvers", "richardson", "ipm", "parallel"):
        return importlib.import_module(f".{name}", __name__)
    raise AttributeError(name)
