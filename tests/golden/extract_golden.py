"""Lift the golden ``z_opt`` vectors and model fixtures out of the reference's own tests.

Run in the build container (needs /root/reference); the output ``reference_golden.json`` is
committed so the tests can run on the GPU box, where the reference tree does not exist.

Sources: tests/test_<F>_<method>.m (``z_opt = [...]``), examples/cl_in_C/main_cl_in_C.c:89-96
(discretised ``[A B]`` to 15 decimals and ``xr`` for ``ur = 0.5``).
"""
import json
import os
import re

REF = os.environ.get('SPCIES_REFERENCE_ROOT', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_golden.json')

TESTS = ['laxMPC_FISTA', 'laxMPC_ADMM', 'equMPC_FISTA', 'equMPC_ADMM', 'ellipMPC_ADMM', 'ellipMPC_ADMM_soc',
         'MPCT_EADMM', 'MPCT_ADMM', 'HMPC_ADMM', 'HMPC_ADMM_s', 'HMPC_SADMM_s']


def main():
    gold = {}
    for t in TESTS:
        path = f'tests/test_{t}.m'
        txt = open(os.path.join(REF, path)).read()
        vecs = re.findall(r'^\s*z_opt = \[(.*?)\];', txt, flags=re.M)
        vec = [v for v in vecs if ';' in v][-1]
        line = txt[:txt.index(vec)].count('\n') + 1
        gold[t] = dict(source=f'{path}:{line}', z_opt=[float(x) for x in vec.split(';')])
    c = open(os.path.join(REF, 'examples/cl_in_C/main_cl_in_C.c')).read()
    ab = re.search(r'AB\[\d*\]\[\d*\]\s*=\s*\{(.*?)\};', c, flags=re.S).group(1)
    rows = re.findall(r'\{([^{}]*)\}', ab)
    gold['main_cl_in_C_AB'] = dict(source='examples/cl_in_C/main_cl_in_C.c:96',
                                   AB=[[float(x) for x in r.split(',')] for r in rows])
    xr = re.search(r'xr\[\w*\]\s*=\s*\{(.*?)\}', c).group(1)
    gold['main_cl_in_C_xr'] = dict(source='examples/cl_in_C/main_cl_in_C.c:89', xr=[float(x) for x in xr.split(',')])
    with open(OUT, 'w') as f:
        json.dump(gold, f, indent=1)
    for k, v in gold.items():
        print(k, v['source'], {kk: (len(vv) if isinstance(vv, list) else vv) for kk, vv in v.items() if kk != 'source'})


if __name__ == '__main__':
    main()
