"""ordinarydiffeq.jl_b200 — B200-native ensemble ODE path behind the reference's
`solve(EnsembleProblem, alg, ensemblealg; ...)` interface.

The directory name contains a dot, so it is imported through `b200_import.load()`
(repo root) which registers it as the module `ordinarydiffeq_jl_b200`.
"""
from . import _lib, codegen, lowlevel, problems_library  # noqa: F401
from ._lib import (ALG_TSIT5, ALG_VERN7, ALG_ROSENBROCK23, ALG_RODAS5P, ALG_DP5, ALG_BS3, ALG_RODAS5, ALG_RODAS4,
                   ALG_RODAS42, ALG_RODAS4P, ALG_RODAS4P2, ALG_VERN6, ALG_VERN8, ALG_VERN9, ALG_ROSENBROCK32, ALG_RODAS5PE, ALG_AUTOTSIT5_ROSENBROCK23, ALG_RODAS3P, ALG_RODAS23W, F32, F64, B200Error, Handle, MultiHandle,  # noqa: F401
                   compile_only)
from .ensemble import (CallbackSet, ContinuousCallback, CSource, DEStats, DiscreteCallback, EnsembleAlgorithm, EnsembleB200, EnsembleContext, EnsembleDistributed,  # noqa: F401
                       EnsembleProblem, EnsembleSerial, EnsembleSolution, EnsembleThreads, ODEFunction, ODEProblem,
                       ODESolution, AutoTsit5, BS3, DP5, Rodas23W, Rodas4, Rodas3P, Rodas42, Rodas4P, Rodas4P2, Rodas5, Rodas5P, Rodas5Pe, Rosenbrock23, Rosenbrock32, TableProbFunc, Tsit5, Vern6, Vern7, Vern8, Vern9, remake, solve)
from . import ranges  # noqa: F401
