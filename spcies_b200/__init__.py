"""spcies_b200 -- B200-native batched solver backend for the Spcies MPC toolbox.

Public surface (mirrors the reference's): :func:`spcies_gen_controller`, :class:`Spcies_options`,
the ``sp_utils`` helpers and the example systems.  The compute path is CUDA only
(``csrc/`` kernels behind the C ABI of ``include/spcies_cuda.h``); there is no CPU fallback.
"""
from .options import Spcies_options, Spcies_problem
from .gen_controller import spcies_gen_controller, make_spec
from . import sp_utils, sysmodel

__version__ = '0.1.0'
__all__ = ['Spcies_options', 'Spcies_problem', 'spcies_gen_controller', 'make_spec', 'sp_utils', 'sysmodel']
