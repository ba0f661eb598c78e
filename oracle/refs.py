"""TEST INFRASTRUCTURE -- not part of the product path.

Named reference builds: for every generated CUDA solver of ``spcies_b200.prebuilt`` the matching
instantiated reference C solver ``oracle/_ref/ref_<save_name>.so`` (see instantiate.py).
"""
from __future__ import annotations

from spcies_b200 import prebuilt

from . import instantiate

_cache = {}


def ref_name(save_name):
    return 'ref_' + save_name


def get(save_name, cflags=('-O3',)):
    key = (save_name, tuple(cflags))
    if key not in _cache:
        spec, cfg = prebuilt.spec_for(save_name)
        tag = ref_name(save_name)
        if tuple(cflags) != ('-O3',):
            tag += '_' + '_'.join(f.strip('-').replace('=', '') for f in cflags)
        _cache[key] = (instantiate.make_reference(spec, tag, cflags), spec, cfg)
    return _cache[key]


N_BOUNDS_VARIANTS = 3


def get_bounds_variant(save_name, s):
    """Reference solver regenerated with bounds variant ``s`` baked in as constants."""
    from spcies_b200 import configs, make_spec
    key = (save_name, 'bounds', s)
    if key not in _cache:
        _, cfg = prebuilt.spec_for(save_name)
        sys2 = configs.bounds_variant(cfg['sys'], s)
        spec = make_spec(sys2, cfg['param'], save_name=f'{save_name}_b{s}', **cfg['kw'])
        _cache[key] = (instantiate.make_reference(spec, f'ref_{save_name}_b{s}'), sys2)
    return _cache[key]


def build_all(verbose=True):
    out = {}
    if not instantiate.reference_available():
        if verbose:
            print('[oracle] reference tree not present: using prebuilt oracle/_ref/*.so')
        return out
    for name in prebuilt.SOLVERS:
        ref, _, _ = get(name)
        out[name] = ref.so_path
        if verbose:
            print(f'[oracle] built {ref.so_path}')
    for s in range(N_BOUNDS_VARIANTS):
        get_bounds_variant('C2_laxMPC_FISTA', s)
    return out
