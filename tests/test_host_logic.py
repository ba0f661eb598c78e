"""CPU tier: host-side mirror of the generator (options, recipes, sparse helpers, projections, emission)."""
import numpy as np
import pytest

from spcies_b200 import Spcies_options, configs, make_spec, sp_utils
from spcies_b200.options import SpciesOptionsError
from spcies_b200.platforms import cuda_code


def test_options_defaults_and_selection():
    o = Spcies_options(formulation='laxMPC')
    assert o.method == 'ADMM' and o.submethod == ''                      # Spcies_options.m:87-106
    assert o.solver['rho'] == 1e-2 and o.solver['k_max'] == 1000          # def_options_laxMPC_ADMM.m
    o = Spcies_options(formulation='MPCT')
    assert o.method == 'EADMM' and o.solver['rho_base'] == 3.0
    o = Spcies_options(formulation='HMPC', method='SADMM')
    assert o.submethod == 'split' and o.solver['alpha'] == 0.95
    o = Spcies_options(formulation='ellipMPC', method='ADMM', submethod='soc', options=dict(rho=15, bogus=3))
    assert o.solver['rho'] == 15 and o.solver['bogus'] == 3               # unknown fields land in .solver (:556-603)
    assert o.platform == 'CUDA' and o.precision == 'double' and o.inf_value == 1e6
    assert o.save_name == 'ellipMPC'


def test_options_validation():
    with pytest.raises(SpciesOptionsError):
        Spcies_options(formulation='nopeMPC')
    with pytest.raises(SpciesOptionsError):
        Spcies_options(formulation='laxMPC', method='ADMM_split')          # the reference's own HMPC tests trip on this
    with pytest.raises(SpciesOptionsError):
        Spcies_options(formulation='laxMPC', platform='Arduino')
    with pytest.raises(SpciesOptionsError):
        Spcies_options(formulation='laxMPC', precision='half')
    assert Spcies_options(formulation='laxMPC', platform='C').platform == 'C'
    assert Spcies_options(formulation='laxMPC', type='equMPC').formulation == 'equMPC'   # deprecated alias


def test_default_defines():
    o = Spcies_options(formulation='laxMPC', method='FISTA', options=dict(debug=False, timing=True))
    names = [r[0] for r in o.default_defCell()]
    assert names == ['MEASURE_TIME', 'in_engineering', 'TIME_VARYING', 'IS_DIAG']       # Spcies_options.m:655-673


def test_sparse_helpers_roundtrip():
    rng = np.random.default_rng(0)
    M = rng.standard_normal((7, 5)) * (rng.random((7, 5)) > 0.5)
    csr = sp_utils.full2CSR(M)
    x = rng.standard_normal(5)
    assert np.allclose(sp_utils.smv(csr.val, csr.col, csr.row, x), M @ x)
    csc = sp_utils.full2CSC(M)
    dense = np.zeros_like(M)
    for c in range(M.shape[1]):
        for j in range(csc.col[c], csc.col[c + 1]):
            dense[csc.row[j], c] = csc.val[j]
    assert np.array_equal(dense, M)
    S = rng.standard_normal((6, 6))
    W = S @ S.T + 6 * np.eye(6)
    val, row, colptr, Dinv = sp_utils.full2LDL(W, True)
    b = rng.standard_normal(6)
    assert np.allclose(sp_utils.LDLsolve(val, row, colptr, Dinv, b), np.linalg.solve(W, b))
    L, D = sp_utils.full2LDL(W)
    assert np.allclose(L @ D @ L.T, W)


def test_projections():
    assert np.allclose(sp_utils.proj_SOC([2.0, 1.0, 1.0]), [2.0, 1.0, 1.0])
    assert np.allclose(sp_utils.proj_SOC([-2.0, 1.0, 1.0]), 0)
    p = sp_utils.proj_SOC([0.0, 3.0, 4.0])
    assert np.isclose(np.linalg.norm(p[1:]), p[0]) and np.allclose(p, [2.5, 1.5, 2.0])
    d = sp_utils.proj_D([5.0, 0.1, 0.1], -1.0, 1.0)                        # diamond: |x1:| <= x0 - lb and <= ub - x0
    assert np.linalg.norm(d[1:]) <= d[0] + 1.0 + 1e-12 and np.linalg.norm(d[1:]) <= 1.0 - d[0] + 1e-12
    x = np.array([0.2, 0.1, -0.05])
    assert np.allclose(sp_utils.proj_D(x, -1.0, 1.0), x)                   # interior point is a fixed point


@pytest.mark.parametrize('name', ['laxMPC_FISTA', 'equMPC_ADMM', 'laxMPC_ADMM', 'ellipMPC_ADMM_soc', 'MPCT_EADMM',
                                  'HMPC_SADMM_split'])
def test_recipe_tables(name):
    cfg = configs.reference_test(name)
    spec = make_spec(cfg['sys'], cfg['param'], **cfg['kw'])
    d = {r.name: r for r in spec.defines}
    assert d['nn_'].value == 6 and d['mm_'].value == 2 and d['NN_'].value == 10
    consts = {r.name: np.asarray(r.value) for r in spec.constants}
    if 'Alpha' in consts:
        assert consts['Alpha'].shape == (9, 6, 6) and consts['Beta'].shape == (10, 6, 6)
        # Beta holds upper-triangular blocks with inverted diagonal: W = Wc' Wc must be reproduced
        W = spec.vars.get('W')
        if W is not None:
            Wc = np.zeros_like(W)
            for i in range(10):
                blk = consts['Beta'][i].copy()
                blk[np.diag_indices(6)] = 1.0 / np.diag(blk)
                Wc[6 * i:6 * i + 6, 6 * i:6 * i + 6] = blk
                if i < 9:
                    Wc[6 * i:6 * i + 6, 6 * i + 6:6 * i + 12] = consts['Alpha'][i]
            assert np.allclose(Wc.T @ Wc, W, atol=1e-10)
    if name == 'ellipMPC_ADMM_soc':
        assert d['dim'].value == 81 and d['n_s'].value == 7 and d['n_eq'].value == 61
        assert consts['L_col'].dtype == np.int32 and consts['L_col'][0] == 0                # 0-based on emission
    if name == 'HMPC_SADMM_split':
        assert d['dim'].value == 98 and d['n_s'].value == 24 and 'alpha_SADMM' in d and 'IS_SYMMETRIC' in d
        assert consts['M1'].shape == (122, 122) and consts['M2'].shape == (122, 6)
    if name == 'MPCT_EADMM':
        assert consts['rho'].shape == (11, 8) and consts['H1i'].shape == (11, 8) and consts['W2'].shape == (8, 8)


def test_cuda_emission_text():
    cfg = configs.reference_test('laxMPC_FISTA')
    spec = make_spec(cfg['sys'], cfg['param'], save_name='emit_check', **cfg['kw'])
    h, cu = cuda_code.emit_text(spec, 'emit_check')
    assert '#define nn_ 6' in h and '#define k_max 5000' in h and '#define tol 0.000000100000000' in h
    assert 'SPCIES_CUDA_DECLARE_SOLVER(laxMPC_FISTA, sol_emit_check);' in h
    assert 'double z[80];' in h and 'double lambda[60];' in h
    assert 'SPCIES_CONST_REAL Alpha[9][6][6];' in cu and '#include "MPC_FISTA.cuh"' in cu
    assert 'R_(-1000.000000000000000)' in cu                                   # %1.15f, dec_var.m:259
    cmd = cuda_code.exec_me('/tmp/emit_check.cu')
    assert 'arch=compute_100a,code=sm_100a' in cmd and '-lineinfo' in cmd
