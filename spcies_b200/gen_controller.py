"""``spcies_gen_controller`` -- entry point of the generator (spcies_gen_controller.m:72-135).

``spcies_gen_controller(sys=..., param=..., formulation=..., method=..., platform='CUDA', ...)``
parses options exactly like the reference (``Spcies_options(varargin{:})``), builds the recipe,
dispatches to the constructor of the selected formulation/method/submethod and hands the
resulting tables to the platform back-end.  Platform ``'CUDA'`` (the one this repository adds)
emits ``<save_name>.cu`` + ``<save_name>.h``, compiles them with ``nvcc`` for sm_100a and
returns a :class:`spcies_b200.solver.CudaSolver`.
"""
from __future__ import annotations

from .formulations import CONSTRUCTORS
from .options import Spcies_options, Spcies_problem


def make_recipe(sys, param, **kw):
    if sys is None:
        raise ValueError("spcies_gen_controller: a 'sys' structure must be provided")
    if param is None:
        raise ValueError("spcies_gen_controller: a 'param' structure must be provided")
    options = kw.pop('spcies_options', None)
    if options is None:
        options = Spcies_options(**kw)
    return Spcies_problem(sys, param, options)


def make_spec(sys=None, param=None, **kw):
    """Recipe -> platform-neutral :class:`SolverSpec` (what ``cons_<F>_<method>_<platform>`` computes
    before any text is emitted)."""
    recipe = make_recipe(sys, param, **kw)
    opt = recipe.options
    if not opt.formulation:
        raise ValueError('Spcies:input_error:no_formulation: The formulation field of options is empty. '
                         'I do not know what to create.')
    key = opt.solver_key()
    if key not in CONSTRUCTORS:
        raise NotImplementedError(f'cons_{key}_{opt.platform}: this formulation/method is not available')
    spec = CONSTRUCTORS[key](recipe)
    spec.options = opt
    from .formulations.common import get_sys_param
    A, B = get_sys_param(recipe)[:2]
    spec.model = (A, B)
    return spec


def spcies_gen_controller(sys=None, param=None, **kw):
    spec = make_spec(sys, param, **kw)
    platform = spec.options.platform
    if platform != 'CUDA':
        raise NotImplementedError(
            f"platform '{platform}' belongs to the reference toolbox (MATLAB); this package emits platform 'CUDA' only")
    from .platforms import cuda_code
    return cuda_code.generate(spec)
