"""Per-formulation recipes (the reference's ``formulations/+<F>/`` packages).

``CONSTRUCTORS`` replaces the reference's dispatch-by-name
``eval(['cons_' F '_' method ['_' sub] '_' platform '(recipe)'])``
(spcies_gen_controller.m:114-130): the key is ``<F>_<method>[_<sub>]`` and the value
builds the platform-neutral :class:`SolverSpec`.
"""
from .common import Row, SolverSpec
from . import laxMPC, ellipMPC, MPCT, HMPC

CONSTRUCTORS = {
    'laxMPC_FISTA': laxMPC.cons_laxMPC_FISTA,
    'laxMPC_ADMM': laxMPC.cons_laxMPC_ADMM,
    'equMPC_FISTA': laxMPC.cons_equMPC_FISTA,
    'equMPC_ADMM': laxMPC.cons_equMPC_ADMM,
    'ellipMPC_ADMM': ellipMPC.cons_ellipMPC_ADMM,
    'ellipMPC_ADMM_soc': ellipMPC.cons_ellipMPC_ADMM_soc,
    'MPCT_EADMM': MPCT.cons_MPCT_EADMM,
    'MPCT_ADMM_cs': MPCT.cons_MPCT_ADMM_cs,
    'MPCT_ADMM_semiband': MPCT.cons_MPCT_ADMM_semiband,
    'HMPC_ADMM': HMPC.cons_HMPC_ADMM,
    'ellipHMPC_ADMM': HMPC.cons_ellipHMPC_ADMM,
    'HMPC_ADMM_split': HMPC.cons_HMPC_ADMM_split,
    'HMPC_SADMM_split': HMPC.cons_HMPC_SADMM_split,
}

__all__ = ['Row', 'SolverSpec', 'CONSTRUCTORS']
