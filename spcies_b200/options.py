"""``Spcies_options`` / ``Spcies_problem`` -- Python mirror of the reference classes.

Follows classes/Spcies_options.m (selection tables :63-106, general defaults
:108-127, constructor :139-272, ``set_opt_from_struct`` :556-603,
``default_defCell`` :655-673) and classes/Spcies_problem.m:24-28.

Differences, all deliberate:

* ``valid_platform`` gains ``'CUDA'`` (the platform this repository adds next to
  ``'C'`` and ``'Matlab'``, Spcies_options.m:65).  ``'CUDA'`` is also the default
  here because it is the only platform this package can emit.
* per-solver defaults (``def_options_<F>_<method>[_<sub>].m``) are a table below
  instead of functions resolved through ``eval`` (Spcies_options.m:494-503).
"""
from __future__ import annotations

import copy as _copy

VALID_FORMULATION = ['laxMPC', 'equMPC', 'ellipMPC', 'MPCT', 'HMPC', 'ellipHMPC', 'personal']
VALID_METHOD = ['ADMM', 'SADMM', 'EADMM', 'FISTA']
VALID_PLATFORM = ['C', 'Matlab', 'CUDA']
VALID_PRECISION = ['double', 'float']

LIST_ACCEPTED_METHODS = {
    'laxMPC': ['ADMM', 'FISTA'], 'equMPC': ['ADMM', 'FISTA'], 'ellipMPC': ['ADMM'],
    'MPCT': ['ADMM', 'EADMM'], 'HMPC': ['ADMM', 'SADMM'], 'ellipHMPC': ['ADMM'],
}
LIST_ACCEPTED_SUBMETHODS = {
    'laxMPC': {'ADMM': [''], 'FISTA': ['']},
    'equMPC': {'ADMM': [''], 'FISTA': ['']},
    'ellipMPC': {'ADMM': ['', 'soc']},
    'MPCT': {'ADMM': ['cs', 'semiband'], 'EADMM': ['']},
    # the reference lists {'cs', 'split'} for HMPC / ADMM (Spcies_options.m:84) but defaults to '' (:104) = cons_HMPC_ADMM_C.m,
    # the non-split solver: the default is accepted here
    'HMPC': {'ADMM': ['', 'cs', 'split'], 'SADMM': ['split']},
    'ellipHMPC': {'ADMM': ['']},
}
LIST_DEF_METHODS = {'laxMPC': 'ADMM', 'equMPC': 'ADMM', 'ellipMPC': 'ADMM',
                    'MPCT': 'EADMM', 'HMPC': 'ADMM', 'ellipHMPC': 'ADMM'}
LIST_DEF_SUBMETHODS = {
    'laxMPC': {'ADMM': '', 'FISTA': ''}, 'equMPC': {'ADMM': '', 'FISTA': ''},
    'ellipMPC': {'ADMM': ''}, 'MPCT': {'ADMM': 'cs', 'EADMM': ''},
    'HMPC': {'ADMM': '', 'SADMM': 'split'}, 'ellipHMPC': {'ADMM': ''},
}

# General defaults, Spcies_options.m:108-127
_BASIC = dict(
    verbose=1, save_name='', directory='$SPCIES$', override=True, const_are_static=True,
    precision='double', inf_value=1e6, save=True, debug=True, timing=True,
    in_engineering=False, time_varying=False, force_diagonal=True,
)

# Solver defaults: formulations/+<F>/def_options_<F>_<method>[_<sub>].m
_HMPC_ADMM = dict(rho=1e-2, sigma=1e-2, tol_p=1e-4, tol_d=1e-4, k_max=1000, box_constraints=None,
                  sparse=False, use_soc=False, alpha=0.95)
_DEF_SOLVER = {
    'laxMPC_FISTA': dict(tol=1e-4, k_max=1000),                                        # def_options_laxMPC_FISTA.m:20-21
    'laxMPC_ADMM': dict(rho=1e-2, tol=1e-4, k_max=1000, force_vector_rho=False),       # def_options_laxMPC_ADMM.m:20-23
    'equMPC_FISTA': dict(tol=1e-4, k_max=1000),
    'equMPC_ADMM': dict(rho=1e-2, tol=1e-4, k_max=1000, force_vector_rho=False),
    'ellipMPC_ADMM': dict(rho=1e-2, tol=1e-4, tol_p=1e-4, tol_d=1e-4, k_max=1000, force_vector_rho=False),
    'ellipMPC_ADMM_soc': dict(rho=5.0, sigma=5.0, tol_p=1e-4, tol_d=1e-4, k_max=1000),
    'MPCT_EADMM': dict(rho_base=3.0, rho_mult=20.0, epsilon_x=1e-6, epsilon_u=1e-6, tol=1e-4, k_max=1000),
    'MPCT_ADMM_cs': dict(rho=1e-2, epsilon_x=1e-6, epsilon_u=1e-6, tol=1e-4, tol_p=1e-4, tol_d=1e-4,
                         k_max=1000, force_vector_rho=False),
    'MPCT_ADMM_semiband': dict(rho=1e-2, epsilon_x=1e-6, epsilon_u=1e-6, epsilon_y=1e-6, tol_p=1e-4, tol_d=1e-4,
                               k_max=1000, force_vector_rho=False, soft_constraints=False,
                               constrained_output=False, beta=1.0),
    # def_options_HMPC_ADMM.m is called for every HMPC submethod ('', cs, split); SADMM delegates to it
    'HMPC_ADMM': dict(_HMPC_ADMM), 'HMPC_ADMM_cs': dict(_HMPC_ADMM), 'HMPC_ADMM_split': dict(_HMPC_ADMM),
    'HMPC_SADMM': dict(_HMPC_ADMM), 'HMPC_SADMM_split': dict(_HMPC_ADMM),
    'ellipHMPC_ADMM': dict(rho=1e-2, tol_p=1e-4, tol_d=1e-4, k_max=1000, box_constraints=None, sparse=False,
                           use_soc=False, alpha=0.95, sigma=0.0),
}

_SELECTION = ('formulation', 'method', 'submethod', 'platform', 'verbose')


class SpciesOptionsError(ValueError):
    pass


class Spcies_options:
    """Options of the toolbox; unknown fields of an options struct go to ``.solver``."""

    def __init__(self, **kw):
        kw = dict(kw)
        options = kw.pop('options', None)
        solver_options = kw.pop('solver_options', None)         # deprecated spelling kept by the reference
        if 'type' in kw:                                        # deprecated alias (Spcies_options.m:199-202)
            kw['formulation'] = kw.pop('type')
        if 'subclass' in kw:
            kw['submethod'] = kw.pop('subclass')
        for d in (options, solver_options):
            if d is not None and not isinstance(d, dict):
                raise SpciesOptionsError('options must be a dict (MATLAB struct)')

        self.solver = {}
        self.to_basic()
        self.formulation = ''
        self.method = ''
        self.submethod = ''
        self.platform = 'CUDA'

        opt = options or {}
        self.formulation = kw.pop('formulation', opt.get('formulation', ''))
        self.platform = kw.pop('platform', opt.get('platform', 'CUDA'))
        self.verbose = kw.pop('verbose', opt.get('verbose', _BASIC['verbose']))
        method = kw.pop('method', opt.get('method', None))
        self.method = method if method is not None else self._default_method()
        sub = kw.pop('submethod', opt.get('submethod', None))
        self.submethod = sub if sub is not None else self._default_submethod()
        self._check_selection()

        self.set_default()
        if options:
            self.set_opt_from_struct(options)
        if solver_options:
            self.set_opt_from_struct(solver_options)
        for name, value in kw.items():
            if name in _BASIC:
                setattr(self, name, value)
            else:
                raise SpciesOptionsError(f"Spcies_options: unknown argument '{name}'")
        if not self.save_name:
            self.save_name = self.formulation

    # -- validated properties (set.* methods, Spcies_options.m:276-397) -------------------
    def __setattr__(self, name, value):
        if name == 'formulation' and value and value not in VALID_FORMULATION:
            raise SpciesOptionsError(f'Spcies_options: formulation {value} is not supported. '
                                     'Please check Spcies_options.valid_formulation')
        if name == 'method' and value and value not in VALID_METHOD:
            raise SpciesOptionsError(f'Spcies_options: method {value} is not supported. '
                                     'Please check Spcies_options.valid_method')
        if name == 'platform' and value not in VALID_PLATFORM:
            raise SpciesOptionsError(f'Spcies_options: platform {value} is not supported. '
                                     'Please check Spcies_options.valid_platform')
        if name == 'precision' and value not in VALID_PRECISION:
            raise SpciesOptionsError(f'Spcies_options: precision {value} is not supported. '
                                     'Please check Spcies_options.valid_precision')
        if name == 'inf_value' and not (isinstance(value, (int, float)) and value > 0):
            raise SpciesOptionsError('Spcies_options: inf_value must be >0')
        if name in ('override', 'const_are_static', 'save', 'debug', 'timing', 'in_engineering',
                    'time_varying', 'force_diagonal') and value not in (True, False, 0, 1):
            raise SpciesOptionsError(f'Spcies_options: {name} must be boolean')
        if name in ('save_name', 'directory') and not isinstance(value, str):
            raise SpciesOptionsError(f'Spcies_options: {name} must be a string of char')
        object.__setattr__(self, name, value)

    # -- defaults ---------------------------------------------------------------------
    def to_basic(self):
        for k, v in _BASIC.items():
            object.__setattr__(self, k, v)
        self.solver = {}

    def _default_method(self):
        if self.formulation and self.formulation != 'personal':
            return LIST_DEF_METHODS[self.formulation]
        return ''

    def _default_submethod(self):
        if self.formulation and self.formulation != 'personal' and self.method:
            try:
                return LIST_DEF_SUBMETHODS[self.formulation][self.method]
            except KeyError:
                raise SpciesOptionsError(f'No default submethod for formulation {self.formulation} '
                                         f'and method {self.method}')
        return ''

    def _check_selection(self):
        f = self.formulation
        if f and f != 'personal':
            if self.method not in LIST_ACCEPTED_METHODS[f]:
                raise SpciesOptionsError(f'method {self.method} is not accepted by formulation {f}')
            if self.submethod not in LIST_ACCEPTED_SUBMETHODS[f][self.method]:
                raise SpciesOptionsError(f'submethod {self.submethod!r} is not accepted by {f} {self.method}')

    def solver_key(self):
        key = self.formulation
        if self.method:
            key += '_' + self.method
        if self.submethod:
            key += '_' + self.submethod
        return key

    def set_default(self):
        defs = _DEF_SOLVER.get(self.solver_key())
        if defs is None:
            self.to_basic()
        else:
            self.set_opt_from_struct(defs)

    def set_opt_from_struct(self, opt, force=True):
        for name, value in opt.items():
            if name in ('formulation', 'method', 'submethod'):
                continue
            if name in _BASIC or name in ('platform',):
                setattr(self, name, value)
            elif force or name in self.solver:
                self.solver[name] = value

    def set(self, name, value):
        self.set_opt_from_struct({name: value}, force=False)

    def force(self, name, value):
        self.set_opt_from_struct({name: value}, force=True)

    def copy(self):
        return _copy.deepcopy(self)

    # -- default #defines, Spcies_options.m:655-673 ---------------------------------------
    def default_defCell(self):
        rows = []
        if self.debug:
            rows.append(('DEBUG', 1, True, 'bool'))
        if self.timing:
            rows.append(('MEASURE_TIME', 1, True, 'bool'))
        rows.append(('in_engineering', int(bool(self.in_engineering)), True, 'int'))
        rows.append(('TIME_VARYING', int(bool(self.time_varying)), True, 'int'))
        if self.force_diagonal:
            rows.append(('IS_DIAG', 1, True, 'bool'))
        return rows


class Spcies_problem:
    """``recipe`` object: controller (sys + param) and options (Spcies_problem.m:24-28)."""

    def __init__(self, sys, param, options: Spcies_options):
        self.sys = sys
        self.param = param
        self.options = options

    def copy(self):
        return Spcies_problem(self.sys, self.param, self.options.copy())
