"""TEST INFRASTRUCTURE -- not part of the product path.

Instantiates the reference's *own* C solver templates, where they lie under
``/root/reference``, exactly the way the reference's MATLAB generator would, and
compiles them with ``gcc -O3`` into ``oracle/_ref/<name>.so``.  Nothing from the
reference is copied into this repository: the instantiated ``.c``/``.h`` live in a
temporary directory that is deleted after compilation; only the binary stays
(``oracle/_ref/`` is git-ignored, but travels to the GPU box).

Python port of the emission half of the reference generator:

* ``dec_var`` / ``write_value``  -- platforms/+C_code/dec_var.m:2-265
* ``declare_variables``          -- platforms/+C_code/declare_variables.m:18-30
* ``construct``                  -- classes/Spcies_constructor.m:95-227 (blocks, snippets :277-321,
                                    $INSERT_NAME$, notice, shared data, final ``fprintf(fid, text)``)
* notice text                    -- +sp_utils/add_notice.m

The numbers substituted into the templates come from the product's recipe tables
(``spcies_b200.formulations``), i.e. the CUDA path and this oracle see the same
``%1.15f`` decimal strings ("Constants are what the C file says").

Only ``tests/``, ``__graft_entry__`` and ``bench.py``'s CPU-baseline leg may import this.
"""
from __future__ import annotations

import ctypes
import hashlib
import json
import os
import re
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get('SPCIES_REFERENCE_ROOT', '/root/reference')
OUT_DIR = os.path.join(HERE, '_ref')
DRIVER_SRC = os.path.join(HERE, 'ref_batch_driver.c')

_C_TYPES = {'float': 'float', 'double': 'double', 'bool': 'boolean', 'int': 'int', 'uint': 'unsigned int',
            'dint': 'long', 'udint': 'unsigned long', 'sint': 'int', 'usint': 'unsigned int'}


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, 'formulations'))


# ---------------------------------------------------------------------------------------
# dec_var.m
# ---------------------------------------------------------------------------------------
def write_value(value, ctype: str) -> str:
    """dec_var.m:241-265: ints ``%d``, bools ``1/0``, reals ``%1.15f``, +-inf -> +-1e20."""
    v = float(value)
    if v == float('inf'):
        v = 1e20
    elif v == float('-inf'):
        v = -1e20
    if 'int' in ctype:
        return '%d' % int(round(v))
    if 'bool' in ctype:
        return '1' if v == 1 else '0'
    return '%1.15f' % v


def dec_var(row) -> str:
    """One declaration line, *before* the final fprintf pass (so it ends in a literal
    backslash-n, dec_var.m:225)."""
    name, value, initialize, typ, options = row.name, row.value, row.initialize, row.type, tuple(row.options)
    if typ not in _C_TYPES:
        raise ValueError('Spcies:dec_var:type_not_recognized')
    ctype = _C_TYPES[typ]
    arr = np.asarray(value)
    if arr.ndim == 0 or arr.size == 1:
        order = 'scalar'
    elif arr.ndim == 1 or (arr.ndim == 2 and min(arr.shape) == 1):
        order = 'vector'
    elif arr.ndim == 2:
        order = 'matrix'
    elif arr.ndim == 3:
        order = '3Dmatrix'
    else:
        raise ValueError(f'Size of variable {name} is not compatible')
    if order == 'scalar' and 'array' in options:
        order = 'vector'
    if 'matrix' in options:
        order = 'matrix'

    out = []
    if 'define' in options:
        if initialize:
            if order == 'scalar':
                out.append('#define %s %s' % (name, write_value(arr.reshape(-1)[0], ctype)))
        else:
            out.append('#define %s' % name)
    else:
        if 'constant' in options:
            out.append('const ')
        if 'static' in options:
            out.append('static ')
        out.append(ctype + ' ' + ('*' if 'pointer' in options else '') + name)
        flat = arr.reshape(-1)
        if order == 'vector':
            out.append('[%d]' % flat.size)
        elif order == 'matrix':
            out.append('[%d][%d]' % (arr.shape[0], arr.shape[1]))
        elif order == '3Dmatrix':
            # numpy arrays are already [block][row][col] == dec_var's [dim3][dim1][dim2]
            out.append('[%d][%d][%d]' % arr.shape)
        if initialize:
            out.append(' = ')
            wv = lambda x: write_value(x, ctype)
            if order == 'scalar':
                out.append(wv(flat[0]))
            elif order == 'vector':
                out.append('{ ' + ', '.join(wv(x) for x in flat) + ' }')
            elif order == 'matrix':
                out.append('{ ' + ', '.join('{' + ', '.join(wv(x) for x in r) + '}' for r in arr) + ' }')
            else:
                out.append('{ ' + ', '.join(
                    '{' + ', '.join('{' + ', '.join(wv(x) for x in r) + '}' for r in blk) + '}' for blk in arr) + ' }')
        out.append(';')
    out.append('\\n')
    return ''.join(out)


def declare_variables(rows):
    text = [dec_var(r) for r in rows]
    return text if text else ['']


# ---------------------------------------------------------------------------------------
# Spcies_constructor.construct
# ---------------------------------------------------------------------------------------
_NOTICE = '\\n\\n// This code is generated by the Spcies toolbox: https://github.com/GepocUS/Spcies\\n'


def _read(rel):
    with open(os.path.join(REF_ROOT, rel), 'r') as f:
        return f.read()


def _insert_snippets(text, ext):
    """Spcies_constructor.m:277-321: every ``spcies_snippet_<name>(...);`` is replaced by
    ``snippets/<name>.<ext>``."""
    for item in re.findall(r'spcies_snippet_(.*?);', text, flags=re.S):
        name = item.split('(')[0]
        text = text.replace('spcies_snippet_' + item + ';', _read(f'snippets/{name}.{ext}'))
    return text


def _fprintf(text):
    """The reference writes files with ``fprintf(fid, text)`` (Spcies_constructor.m:222), i.e. the
    text is a *format string*: ``%%`` -> ``%`` and backslash escapes are interpreted."""
    out, i, n = [], 0, len(text)
    esc = {'n': '\n', 't': '\t', '\\': '\\', 'r': '\r', 'a': '\a', 'b': '\b', 'f': '\f', 'v': '\v'}
    while i < n:
        c = text[i]
        if c == '%' and i + 1 < n and text[i + 1] == '%':
            out.append('%')
            i += 2
        elif c == '\\' and i + 1 < n and text[i + 1] in esc:
            out.append(esc[text[i + 1]])
            i += 2
        else:
            out.append(c)
            i += 1
    return ''.join(out)


# Defects of the reference's HEADER templates that stop an option from compiling at all; the prototype is corrected in the
# instantiated text (the .c template -- the algorithm -- is never touched):
#   header_laxMPC_ADMM_C.h:25  the TIME_VARYING prototype names the solution type `solution` instead of sol_$INSERT_NAME$
#                              (the definition in code_laxMPC_ADMM_C.c:19 has the right type)
REF_HEADER_FIXUPS = {
    'formulations/+laxMPC/header_laxMPC_ADMM_C.h': [('int *e_flag, solution *sol);', 'int *e_flag, sol_$INSERT_NAME$ *sol);')],
}


def construct(spec, save_name):
    """Return ``{'c': text, 'h': text}`` of the files the reference would write for platform 'C'."""
    files = {
        'c': _read('platforms/+C_code/generic_solver_struct.c').replace('$INSERT_SOLVER$', _read(spec.ref_code)),
        'h': _read(spec.ref_header),
    }
    for old_, new_ in REF_HEADER_FIXUPS.get(spec.ref_header, ()):
        files['h'] = files['h'].replace(old_, new_)
    for ext in files:
        t = _insert_snippets(files[ext], ext)
        t = t.replace('$INSERT_NAME$', save_name).replace('$INSERT_PATH$', '')
        files[ext] = t + _NOTICE
    shared = [('$INSERT_DEFINES$', spec.defines), ('$INSERT_CONSTANTS$', spec.constants)]
    if spec.variables or '$INSERT_VARIABLES$' in files['c']:
        shared.append(('$INSERT_VARIABLES$', spec.variables))
    for tag, rows in shared:
        strings = declare_variables(rows)
        block = ''.join(strings)        # equivalent to the k-loop of strrep(tag, [string_k tag]) + final strrep
        for ext in files:
            files[ext] = files[ext].replace(tag, block)
    return {ext: _fprintf(t) for ext, t in files.items()}


# ---------------------------------------------------------------------------------------
# Build + ctypes wrapper
# ---------------------------------------------------------------------------------------
def _spec_digest(spec, save_name, flags):
    h = hashlib.sha256()
    h.update(save_name.encode())
    h.update(' '.join(flags).encode())
    for rows in (spec.defines, spec.constants, spec.variables):
        for r in rows:
            h.update(dec_var(r).encode())
    with open(DRIVER_SRC, 'rb') as f:
        h.update(f.read())
    return h.hexdigest()[:16]


def build_reference(spec, save_name, cflags=('-O3',), force=False):
    """Instantiate + compile; returns the path of ``oracle/_ref/<save_name>.so``.

    If the reference tree is absent (GPU box) the prebuilt file is used as is.
    """
    os.makedirs(OUT_DIR, exist_ok=True)
    so_path = os.path.join(OUT_DIR, save_name + '.so')
    stamp = os.path.join(OUT_DIR, save_name + '.stamp')
    if not reference_available():
        if os.path.exists(so_path):
            return so_path
        raise FileNotFoundError(f'{so_path} was not prebuilt and the reference tree is not available')
    digest = _spec_digest(spec, save_name, list(cflags))
    if not force and os.path.exists(so_path) and os.path.exists(stamp) and open(stamp).read() == digest:
        return so_path
    files = construct(spec, save_name)
    tmp = tempfile.mkdtemp(prefix='spcies_ref_')
    try:
        with open(os.path.join(tmp, save_name + '.c'), 'w') as f:
            f.write(files['c'])
        with open(os.path.join(tmp, save_name + '.h'), 'w') as f:
            f.write(files['h'])
        has_r = 1 if 'r_ellip' in spec.extra_inputs else 0
        nref = 3 if 'xrs' in spec.extra_inputs else 1
        cmd = ['gcc', *cflags, '-fPIC', '-shared', '-w', '-I', tmp, f'-DSPCIES_NREF={nref}',
               f'-DSPCIES_HDR="{save_name}.h"', f'-DSPCIES_FUNC={spec.func_name}',
               f'-DSPCIES_SOL=sol_{save_name}', f'-DSPCIES_HAS_R={has_r}',
               os.path.join(tmp, save_name + '.c'), DRIVER_SRC, '-o', so_path, '-lm', '-lpthread']
        subprocess.run(cmd, check=True, capture_output=True, text=True)
        meta = _sol_layout(files['h'], save_name)
        with open(os.path.join(OUT_DIR, save_name + '.layout.json'), 'w') as f:
            json.dump(meta, f)
        with open(stamp, 'w') as f:
            f.write(digest)
    except subprocess.CalledProcessError as e:
        raise RuntimeError('gcc failed on the instantiated reference template:\n' + e.stderr[-4000:])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return so_path


def _sol_layout(header_text, save_name):
    """Parse ``typedef struct {...} sol_<name>;`` -> [(field, length)] with lengths evaluated
    from the ``#define``s of the same header."""
    defs = {}
    for mline in re.finditer(r'^#define\s+(\w+)\s+([-\d.eE+]+)\s*$', header_text, flags=re.M):
        try:
            defs[mline.group(1)] = int(float(mline.group(2)))
        except ValueError:
            pass
    body = re.search(r'typedef struct\s*\{(.*?)\}\s*sol_' + re.escape(save_name) + r'\s*;', header_text, flags=re.S)
    fields = []
    seen = set()
    for fm in re.finditer(r'double\s+(\w+)\s*(?:\[(.*?)\])?\s*;', body.group(1)):
        if fm.group(1) in seen:          # the other branch of an #if inside the struct (header_MPCT_ADMM_semiband_C.h:16-22)
            continue
        try:
            length = 1 if fm.group(2) is None else int(eval(fm.group(2), {'__builtins__': {}}, defs))
        except NameError:
            continue
        seen.add(fm.group(1))
        fields.append((fm.group(1), length))
    return fields


class RefSolver:
    """ctypes handle on an instantiated + compiled reference solver."""

    def __init__(self, spec, save_name, so_path=None):
        self.spec = spec
        self.save_name = save_name
        self.so_path = so_path or os.path.join(OUT_DIR, save_name + '.so')
        self.lib = ctypes.CDLL(self.so_path)
        with open(os.path.join(OUT_DIR, save_name + '.layout.json')) as f:
            self.layout = [(str(a), int(b)) for a, b in json.load(f)]
        self.n, self.m = spec.dims['n'], spec.dims['m']
        self.has_r = 'r_ellip' in spec.extra_inputs
        self.nref = 3 if 'xrs' in spec.extra_inputs else 1
        self.tv = 'A_in' in spec.extra_inputs            # TIME_VARYING: per-instance model
        self.sol_len = sum(l for _, l in self.layout)
        self.lib.spcies_ref_sol_doubles.restype = ctypes.c_long
        assert self.lib.spcies_ref_sol_doubles() == self.sol_len, 'sol_<name> layout mismatch'
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        self.lib.spcies_ref_batch.restype = ctypes.c_int
        self.lib.spcies_ref_batch.argtypes = [ctypes.c_long, dp, dp, dp, dp, dp, ip, ip, dp, ctypes.c_int]

    @staticmethod
    def _p(a, t=ctypes.c_double):
        return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))

    def solve_batch(self, x0, xr, ur, r=None, threads=1, want_sol=False, tv=None, LB=None, UB=None):
        """Loop of single-instance reference calls (in C).  Returns ``u [B,m], k [B], e [B]`` (+ sol dict).
        TIME_VARYING solvers: ``tv = (A [B,n,n], Bm [B,n,m], Q [B,n], R [B,m])`` (row-major matrices; passed column-major to the
        reference, as MATLAB does) and ``LB / UB [B, nm]``."""
        x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
        B = x0.shape[0]
        keep = []
        if self.tv:
            A_, B_, Q_, R_ = tv
            Ac = np.ascontiguousarray(np.transpose(np.asarray(A_, float).reshape(B, self.n, self.n), (0, 2, 1)))
            Bc = np.ascontiguousarray(np.transpose(np.asarray(B_, float).reshape(B, self.n, self.m), (0, 2, 1)))
            Qc = np.ascontiguousarray(np.asarray(Q_, float).reshape(B, self.n))
            Rc = np.ascontiguousarray(np.asarray(R_, float).reshape(B, self.m))
            LBc = np.ascontiguousarray(np.asarray(LB, float).reshape(B, self.n + self.m))
            UBc = np.ascontiguousarray(np.asarray(UB, float).reshape(B, self.n + self.m))
            keep += [Ac, Bc, Qc, Rc, LBc, UBc]
            self.lib.spcies_ref_set_refs(self._p(Ac), self._p(Bc), self._p(Qc), self._p(Rc))
            self.lib.spcies_ref_set_bounds(self._p(LBc), self._p(UBc))
        if self.nref == 3:      # xr = (x_re, x_rs, x_rc), ur = (u_re, u_rs, u_rc)
            xs = [np.ascontiguousarray(np.atleast_2d(a), dtype=np.float64) for a in xr]
            us = [np.ascontiguousarray(np.atleast_2d(a), dtype=np.float64) for a in ur]
            assert all(a.shape == (B, self.n) for a in xs) and all(a.shape == (B, self.m) for a in us)
            xr, ur = xs[0], us[0]
            self.lib.spcies_ref_set_refs(self._p(xs[1]), self._p(xs[2]), self._p(us[1]), self._p(us[2]))
        xr = np.ascontiguousarray(np.atleast_2d(xr), dtype=np.float64)
        ur = np.ascontiguousarray(np.atleast_2d(ur), dtype=np.float64)
        assert x0.shape == (B, self.n) and xr.shape == (B, self.n) and ur.shape == (B, self.m)
        if self.has_r:
            r = np.ascontiguousarray(np.broadcast_to(np.asarray(r, dtype=np.float64).ravel(), (B,)))
        u = np.empty((B, self.m))
        k = np.empty(B, dtype=np.int32)
        e = np.empty(B, dtype=np.int32)
        sol = np.zeros((B, self.sol_len)) if want_sol else None
        rc = self.lib.spcies_ref_batch(B, self._p(x0), self._p(xr), self._p(ur), self._p(r if self.has_r else None),
                                       self._p(u), self._p(k, ctypes.c_int), self._p(e, ctypes.c_int),
                                       self._p(sol), int(threads))
        if rc != 0:
            raise RuntimeError(f'spcies_ref_batch failed ({rc})')
        if not want_sol:
            return u, k, e
        out, off = {}, 0
        for name, length in self.layout:
            out[name] = sol[:, off:off + length]
            off += length
        return u, k, e, out

    def closed_loop(self, x0, xr, ur, steps, AB, r=None, threads=1):
        """Loop of single-instance reference calls with the plant ``x+ = AB (x; u)`` in between (in C, the example's operation
        order).  Returns ``x [steps+1, B, n], u [steps, B, m], k [steps, B], e [steps, B]``."""
        x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
        xr = np.ascontiguousarray(np.atleast_2d(xr), dtype=np.float64)
        ur = np.ascontiguousarray(np.atleast_2d(ur), dtype=np.float64)
        AB = np.ascontiguousarray(AB, dtype=np.float64)
        B = x0.shape[0]
        assert AB.shape == (self.n, self.n + self.m)
        if self.has_r:
            r = np.ascontiguousarray(np.broadcast_to(np.asarray(r, dtype=np.float64).ravel(), (B,)))
        xt = np.empty((steps + 1, B, self.n))
        ut = np.empty((steps, B, self.m))
        kt = np.empty((steps, B), dtype=np.int32)
        et = np.empty((steps, B), dtype=np.int32)
        dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
        f = self.lib.spcies_ref_closed_loop
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_long, ctypes.c_int, dp, dp, dp, dp, dp, dp, dp, ip, ip, ctypes.c_int]
        rc = f(B, int(steps), self._p(x0), self._p(xr), self._p(ur), self._p(r if self.has_r else None), self._p(AB),
               self._p(xt), self._p(ut), self._p(kt, ctypes.c_int), self._p(et, ctypes.c_int), int(threads))
        if rc != 0:
            raise RuntimeError(f'spcies_ref_closed_loop failed ({rc})')
        return xt, ut, kt, et

    def solve(self, x0, xr, ur, r=None):
        if self.nref == 3:
            xr, ur = [np.asarray(a)[None] for a in xr], [np.asarray(a)[None] for a in ur]
        else:
            xr, ur = np.asarray(xr)[None], np.asarray(ur)[None]
        u, k, e, sol = self.solve_batch(np.asarray(x0)[None], xr, ur,
                                        r=None if r is None else np.asarray([r], dtype=float), want_sol=True)
        return u[0], int(k[0]), int(e[0]), {name: v[0] for name, v in sol.items()}


def make_reference(spec, save_name, cflags=('-O3',)):
    return RefSolver(spec, save_name, build_reference(spec, save_name, cflags))
