"""Platform ``'CUDA'``: emission and compilation of a generated solver for sm_100a.

The counterpart, for this platform, of the reference's ``platforms/+C_code`` package
(dec_var.m:2-265, declare_variables.m, generic_solver_struct.c) plus the file-writing half of
classes/Spcies_constructor.m:95-251.  For a recipe it writes, into ``options.directory``:

* ``<save_name>.h``  -- the reference's ``#define`` block (same names, same ``%1.15f`` text), the
  ``sol_<save_name>`` struct and the prototypes of the single-instance and batched entry points;
* ``<save_name>.cu`` -- the constants as one POD ``struct spcies_consts`` (same member names and
  ``[block][row][col]`` layout as the reference's ``const static`` arrays) and an ``#include`` of
  the hand-written kernel template under ``spcies_b200/csrc/``;

and runs the platform's ``exec_me`` (``nvcc -gencode arch=compute_100a,code=sm_100a ...``, the
analogue of ``mex ... COPTIMFLAGS="-O3"``, +sp_utils/get_generic_mex_exec.m:20-31) to build
``<save_name>.so``, the C-ABI library of ``include/spcies_cuda.h``.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC_DIR = os.path.join(PKG_DIR, 'csrc')
REPO_DIR = os.path.dirname(PKG_DIR)
INCLUDE_DIR = os.path.join(REPO_DIR, 'include')
DEFAULT_DIR = os.path.join(REPO_DIR, 'generated_solvers')

NVCC_ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
NVCC_FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
              '-Xcompiler', '-fno-gnu-unique', '-shared', '-Xptxas', '-v']

# kernel template per solver + compile-time switches it needs
KERNELS = {
    'laxMPC_FISTA': ('MPC_FISTA.cuh', {'SPCIES_TERMINAL': 1}),
    'equMPC_FISTA': ('MPC_FISTA.cuh', {'SPCIES_TERMINAL': 0}),
    'laxMPC_ADMM': ('MPC_ADMM.cuh', {'SPCIES_TERMINAL': 1}),
    'equMPC_ADMM': ('MPC_ADMM.cuh', {'SPCIES_TERMINAL': 0}),
    'ellipMPC_ADMM': ('MPC_ADMM.cuh', {'SPCIES_TERMINAL': 2}),
    'ellipMPC_ADMM_soc': ('ellipMPC_ADMM_soc.cuh', {}),
    'MPCT_EADMM': ('MPCT_EADMM.cuh', {}),
    'MPCT_ADMM_cs': ('MPCT_ADMM_cs.cuh', {}),
    'MPCT_ADMM_semiband': ('MPCT_ADMM_semiband.cuh', {}),
    'HMPC_ADMM_split': ('HMPC_ADMM_split.cuh', {}),
    'HMPC_ADMM': ('HMPC_ADMM.cuh', {'SPCIES_NREF': 1}),
    'ellipHMPC_ADMM': ('HMPC_ADMM.cuh', {'SPCIES_NREF': 3}),
}


# solvers generated with options.time_varying (per-instance model, factorisation on the device)
KERNELS_TV = {
    'laxMPC_FISTA': ('MPC_FISTA_tv.cuh', {'SPCIES_TERMINAL': 1}),
    'equMPC_FISTA': ('MPC_FISTA_tv.cuh', {'SPCIES_TERMINAL': 0}),
    'equMPC_ADMM': ('MPC_ADMM_tv.cuh', {'SPCIES_TERMINAL': 0}),
    'laxMPC_ADMM': ('MPC_ADMM_tv.cuh', {'SPCIES_TERMINAL': 1}),
}


def write_value(value, is_int=False, is_bool=False):
    """Number formats of dec_var.m:241-265 (ints ``%d``, reals ``%1.15f``, +-inf -> +-1e20)."""
    v = float(value)
    if v == float('inf'):
        v = 1e20
    elif v == float('-inf'):
        v = -1e20
    if is_int:
        return '%d' % int(round(v))
    if is_bool:
        return '1' if v == 1 else '0'
    return '%1.15f' % v


def _is_int_type(t):
    return t in ('int', 'uint', 'dint', 'udint', 'sint', 'usint')


def _define_line(row):
    if not row.initialize:
        return f'#define {row.name}'
    v = np.asarray(row.value).reshape(-1)[0]
    return f'#define {row.name} ' + write_value(v, _is_int_type(row.type), row.type == 'bool')


def _member_decl(row):
    arr = np.asarray(row.value)
    ctype = 'int' if _is_int_type(row.type) else 'SPCIES_CONST_REAL'
    if arr.ndim == 0:
        return f'    {ctype} {row.name};'
    dims = ''.join('[%d]' % d for d in arr.shape)
    return f'    {ctype} {row.name}{dims};'


def _member_init(row):
    arr = np.asarray(row.value)
    is_int = _is_int_type(row.type)
    fmt = (lambda x: write_value(x, True)) if is_int else (lambda x: 'R_(' + write_value(x) + ')')

    def rec(a):
        if a.ndim == 0:
            return fmt(a)
        if a.ndim == 1:
            return '{' + ', '.join(fmt(x) for x in a) + '}'
        return '{' + ', '.join(rec(x) for x in a) + '}'
    return '    /* %s */ %s' % (row.name, rec(arr))


def emit_text(spec, save_name):
    """Return ``(header_text, cu_text)`` for a recipe."""
    opts = spec.options
    key = spec.kernel
    if key not in KERNELS:
        raise NotImplementedError(f'no CUDA kernel template for {key}')
    kernel_hdr, switches = KERNELS[key]
    if opts.time_varying:
        if key not in KERNELS_TV:
            raise NotImplementedError(f'time_varying: no CUDA kernel template for {key} (available: {sorted(KERNELS_TV)})')
        kernel_hdr, switches = KERNELS_TV[key]
    sol_t = f'sol_{save_name}'
    has_r = 1 if 'r_ellip' in spec.extra_inputs else 0

    h = [f'#ifndef {save_name}_h', f'#define {save_name}_h', '']
    h += [_define_line(r) for r in spec.defines]
    h += ['', '#include "spcies_cuda.h"', '', 'typedef struct {']
    for fname, length in spec.sol_fields:
        h.append(f'    double {fname}[{length}];')
    h += ['    double update_time; // host->device copy of the inputs (ms)',
          '    double solve_time; // solver kernel on the device (ms)',
          '    double polish_time; // device->host copy of the results (ms)',
          '    double run_time; // whole call (ms)',
          '} ' + sol_t + ';', '', '#ifdef __cplusplus', 'extern "C" {', '#endif']
    macro = 'SPCIES_CUDA_DECLARE_SOLVER_R' if has_r else 'SPCIES_CUDA_DECLARE_SOLVER'
    if 'xrs' in spec.extra_inputs:
        macro = 'SPCIES_CUDA_DECLARE_SOLVER_6REF'
    if opts.time_varying:
        macro = 'SPCIES_CUDA_DECLARE_SOLVER_TV'
    h += [f'{macro}({spec.func_name}, {sol_t});', '#ifdef __cplusplus', '}', '#endif', '', '#endif', '',
          '// This code is generated by the CUDA platform of spcies_b200 for the Spcies toolbox: '
          'https://github.com/GepocUS/Spcies', '']

    # precision = 'float': platforms/+C_code/dec_var.m:16-17 only changes the *declared type of the emitted constants*; the
    # locals of every template stay `double`, i.e. the reference's float solver computes in double on float-rounded constants.
    # The CUDA platform does the same by default (constants float, arithmetic double: every tensor-core engine applies and the
    # results match the float-generated reference like the double ones do).  options.solver['float_arithmetic'] = True selects
    # true single-precision arithmetic instead (one-thread-per-instance kernels; gate 1e-5).
    const_real = 'float' if opts.precision == 'float' else 'double'
    real = 'float' if (opts.precision == 'float' and opts.solver.get('float_arithmetic', False)) else 'double'
    cu = [f'// {save_name}.cu -- generated solver: {spec.formulation} {spec.method} {spec.submethod}'.rstrip(),
          f'#include "{save_name}.h"', '',
          f'#define SPCIES_REAL {real}',
          f'#define SPCIES_CONST_REAL {const_real}',
          f'#define SPCIES_PRECISION_STR "{const_real}"',
          f'#define SPCIES_ARITH_STR "{real}"',
          f'#define SPCIES_FUNC {spec.func_name}',
          f'#define SPCIES_SOL_T {sol_t}',
          f'#define SPCIES_SOLVER_STR "{opts.solver_key()}"',
          f'#define SPCIES_SAVE_NAME_STR "{save_name}"',
          f'#define SPCIES_HAS_R {has_r}']
    cu += [f'#define {k} {v}' for k, v in switches.items()]
    rows = [r for r in list(spec.constants) + list(spec.variables)]
    cu += ['#define R_(x) ((SPCIES_CONST_REAL)(x))', '', 'struct alignas(16) spcies_consts {']
    cu += [_member_decl(r) for r in rows]
    cu += ['};', '', 'static const spcies_consts spcies_h_consts = {']
    cu.append(',\n'.join(_member_init(r) for r in rows))
    # the prediction model [A B], row-major: the default plant of <func>_closed_loop (exact round trip: 17 significant digits)
    AB = np.hstack([np.asarray(spec.model[0], float), np.asarray(spec.model[1], float)])
    cu += ['};', '', 'static const double spcies_model_AB[nn_ * nm_] = {' + ', '.join('%.17g' % v for v in AB.ravel()) + '};']
    cu += ['', f'#include "{kernel_hdr}"', '',
           '// This code is generated by the CUDA platform of spcies_b200 for the Spcies toolbox: '
           'https://github.com/GepocUS/Spcies', '']
    return '\n'.join(h), '\n'.join(cu)


def resolve_directory(opts):
    d = opts.directory
    if d in ('', '$SPCIES$'):
        d = DEFAULT_DIR
    return d


def emit(spec, directory=None, save_name=None):
    opts = spec.options
    directory = directory or resolve_directory(opts)
    save_name = save_name or opts.save_name
    os.makedirs(directory, exist_ok=True)
    h_text, cu_text = emit_text(spec, save_name)
    h_path = os.path.join(directory, save_name + '.h')
    cu_path = os.path.join(directory, save_name + '.cu')
    for path, text in ((h_path, h_text), (cu_path, cu_text)):
        if not (os.path.exists(path) and open(path).read() == text):
            with open(path, 'w') as f:
                f.write(text)
    return cu_path, h_path


def nvcc_path():
    p = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    return p if os.path.exists(p) else None


def _included_headers(path, seen=None):
    """The .cuh / .h files under csrc/ that ``path`` includes, transitively (so that a new kernel header does not invalidate
    the build stamp of every other generated solver)."""
    import re
    seen = set() if seen is None else seen
    try:
        text = open(path).read()
    except OSError:
        return seen
    for inc in re.findall(r'^\s*#\s*include\s+"([^"]+)"', text, flags=re.M):
        q = os.path.join(CSRC_DIR, inc)
        if os.path.exists(q) and q not in seen:
            seen.add(q)
            _included_headers(q, seen)
    return seen


def _sources_digest(cu_path, extra_flags):
    h = hashlib.sha256()
    h.update(' '.join(NVCC_ARCH + NVCC_FLAGS + list(extra_flags)).encode())
    for p in [cu_path, cu_path[:-3] + '.h', os.path.join(INCLUDE_DIR, 'spcies_cuda.h')] + sorted(_included_headers(cu_path)):
        with open(p, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def exec_me(cu_path, extra_flags=()):
    """The command the constructor runs after writing the files (``exec_me`` of the platform)."""
    so_path = cu_path[:-3] + '.so'
    return [nvcc_path() or 'nvcc', *NVCC_ARCH, *NVCC_FLAGS, *extra_flags,
            '-I', INCLUDE_DIR, '-I', CSRC_DIR, '-I', os.path.dirname(cu_path), '-o', so_path, cu_path]


def build(cu_path, extra_flags=(), force=False, verbose=False):
    """Compile ``<name>.cu`` -> ``<name>.so`` (skipped when sources and flags are unchanged).

    When nvcc is unavailable (it is present in the build container and on the GPU box) a prebuilt,
    up-to-date ``.so`` is accepted; otherwise this raises -- there is no fallback path.
    """
    so_path = cu_path[:-3] + '.so'
    stamp = cu_path[:-3] + '.stamp'
    digest = _sources_digest(cu_path, extra_flags)
    if not force and os.path.exists(so_path) and os.path.exists(stamp) and open(stamp).read().split('\n')[0] == digest:
        return so_path
    if nvcc_path() is None:
        raise RuntimeError('nvcc not found: cannot build ' + so_path)
    cmd = exec_me(cu_path, extra_flags)
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + (p.stdout + p.stderr)[-6000:])
    log = p.stdout + p.stderr
    with open(stamp, 'w') as f:
        f.write(digest + '\n' + log)
    if verbose:
        print(log)
    return so_path


def generate(spec, directory=None, save_name=None, extra_flags=(), force=False):
    """emit + build + load: what ``constructor.construct(options)`` does for this platform."""
    from ..solver import CudaSolver
    cu_path, h_path = emit(spec, directory, save_name)
    so_path = build(cu_path, extra_flags=extra_flags, force=force)
    return CudaSolver(so_path, spec)
