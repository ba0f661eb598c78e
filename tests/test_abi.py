"""CPU tier: every generated library loads and exports the symbols include/spcies_cuda.h declares for its solver
family (no compute calls here -- there is no GPU), and refuses to run without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from spcies_b200 import prebuilt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, 'include', 'spcies_cuda.h')).read()

COMMON = re.findall(r'^\s*(?:int|long|void|const char \*)\s*\*?(spcies_cuda_\w+)\(', HEADER, flags=re.M)
FAMILIES = dict(re.findall(r'SPCIES_CUDA_SOLVER(?:_R)?\((\w+)\)\s+(\w+)\s+\w+_batch', HEADER))


def test_header_lists_common_symbols_and_families():
    assert set(COMMON) >= {'spcies_cuda_abi_version', 'spcies_cuda_solver_name', 'spcies_cuda_save_name',
                           'spcies_cuda_precision', 'spcies_cuda_dims', 'spcies_cuda_sol_doubles',
                           'spcies_cuda_device_count', 'spcies_cuda_kernel_attributes', 'spcies_cuda_free',
                           'spcies_cuda_last_error'}
    assert set(FAMILIES) == {'laxMPC_FISTA', 'laxMPC_ADMM', 'equMPC_FISTA', 'equMPC_ADMM', 'ellipMPC_ADMM',
                             'ellipMPC_ADMM_soc', 'MPCT_EADMM', 'MPCT_ADMM_cs', 'MPCT_ADMM_semiband', 'HMPC_ADMM'}


@pytest.mark.parametrize('name', list(prebuilt.SOLVERS))
def test_library_exports_declared_symbols(name):
    spec, _cfg = prebuilt.spec_for(name)
    so = os.path.join(ROOT, 'generated_solvers', name + '.so')
    assert os.path.exists(so), 'run __graft_entry__.build() first'
    lib = ctypes.CDLL(so)
    for sym in COMMON:
        assert hasattr(lib, sym), sym
    assert spec.func_name in FAMILIES
    assert hasattr(lib, spec.func_name) and hasattr(lib, spec.func_name + '_batch')
    lib.spcies_cuda_solver_name.restype = ctypes.c_char_p
    lib.spcies_cuda_save_name.restype = ctypes.c_char_p
    assert lib.spcies_cuda_save_name().decode() == name
    assert lib.spcies_cuda_solver_name().decode() == spec.options.solver_key()
    assert lib.spcies_cuda_abi_version() == 2
    if 'xrs' not in spec.extra_inputs and 'A_in' not in spec.extra_inputs:
        assert hasattr(lib, spec.func_name + '_closed_loop')
    nn, mm, NN = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib.spcies_cuda_dims(ctypes.byref(nn), ctypes.byref(mm), ctypes.byref(NN))
    assert (nn.value, mm.value, NN.value) == (spec.dims['n'], spec.dims['m'], spec.dims['N'])
    lib.spcies_cuda_sol_doubles.restype = ctypes.c_long
    assert lib.spcies_cuda_sol_doubles() == sum(l for _, l in spec.sol_fields) + 4


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the batched entry returns SPCIES_CUDA_ENODEVICE (the product path never falls back)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from spcies_b200.solver import CudaSolver, SpciesCudaError
    spec, cfg = prebuilt.spec_for('C2_laxMPC_FISTA')
    sol = CudaSolver(os.path.join(ROOT, 'generated_solvers', 'C2_laxMPC_FISTA.so'), spec)
    assert sol.device_count() == 0
    x0 = np.zeros((4, 6)); xr = np.zeros((4, 6)); ur = np.zeros((4, 2))
    with pytest.raises(SpciesCudaError, match='1002'):
        sol.solve_batch(x0, xr, ur)
    with pytest.raises(SpciesCudaError):
        sol.solve(x0[0], xr[0], ur[0])


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The Python mirror of spcies_batch_opts / spcies_batch_info (spcies_b200/solver.py) has the size and field offsets of
    the C structs in include/spcies_cuda.h (compiled here with gcc: the header is plain C)."""
    import ctypes
    import os
    import subprocess
    from spcies_b200.solver import BatchInfo, BatchOpts
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / 'sz.c'
    fields_o = ['device', 'n_devices', 'arith', 'device_pointers', 'LB', 'UB', 'stream', 'block_threads', 'grid_blocks', 'tail_mode',
                'tail_grace', 'engine', 'tail_caps', 'warm_start', 'reserved', 'plant_AB']
    fields_i = ['kernel_ms', 'launches', 'sum_k', 'block_threads', 'n_devices', 'drain_us', 'parked', 'reserved']
    body = ''.join(f'printf("o {f} %zu\\n", offsetof(spcies_batch_opts, {f}));\n' for f in fields_o)
    body += ''.join(f'printf("i {f} %zu\\n", offsetof(spcies_batch_info, {f}));\n' for f in fields_i)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "spcies_cuda.h"\nint main(void) {\n'
                   'printf("o sizeof %zu\\n", sizeof(spcies_batch_opts)); printf("i sizeof %zu\\n", sizeof(spcies_batch_info));\n'
                   + body + 'return 0; }\n')
    exe = tmp_path / 'sz'
    subprocess.run(['gcc', '-I', os.path.join(root, 'include'), str(src), '-o', str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    for line in out.splitlines():
        which, name, val = line.split()
        cls = BatchOpts if which == 'o' else BatchInfo
        expect = ctypes.sizeof(cls) if name == 'sizeof' else getattr(cls, name).offset
        assert int(val) == expect, (which, name, val, expect)
