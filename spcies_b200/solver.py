"""ctypes binding of a generated solver library (the C ABI of ``include/spcies_cuda.h``).

``CudaSolver`` is what ``spcies_gen_controller(..., platform='CUDA')`` returns.  Its call
signatures mirror the reference's MEX gateways (``[u, k, e_flag, sol] = name(x0, xr, ur)``,
formulations/+laxMPC/struct_laxMPC_FISTA_C_Matlab.c:8-167):

* ``solve(x0, xr, ur[, r])``        -> ``u, k, e_flag, sol``  through the *unchanged* single-instance symbol
* ``solve_batch(x0, xr, ur[, r])``  -> ``u [B,m], k [B], e_flag [B], info``  through ``<func>_batch``

There is no CPU path: if the library cannot run on a CUDA device the calls raise.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_char_p, c_double, c_int, c_long, c_void_p

import numpy as np


class BatchOpts(ctypes.Structure):
    _fields_ = [('device', c_int), ('n_devices', c_int), ('arith', c_int), ('device_pointers', c_int),
                ('LB', c_void_p), ('UB', c_void_p), ('stream', c_void_p),
                ('block_threads', c_int), ('grid_blocks', c_int), ('tail_mode', c_int), ('tail_grace', c_int),
                ('engine', c_int), ('tail_caps', c_int * 3), ('warm_start', c_int), ('reserved', c_int * 1), ('plant_AB', c_void_p)]


class BatchInfo(ctypes.Structure):
    _fields_ = [('kernel_ms', c_double), ('h2d_ms', c_double), ('d2h_ms', c_double), ('total_ms', c_double),
                ('launches', c_long), ('h2d_bytes', c_long), ('d2h_bytes', c_long), ('sum_k', c_long),
                ('n_not_converged', c_long),
                ('block_threads', c_int), ('grid_blocks', c_int), ('smem_bytes', c_int), ('regs_per_thread', c_int),
                ('n_devices', c_int), ('drain_us', c_int), ('span_us', c_int), ('parked', c_int), ('reserved', c_int * 4)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_ if f != 'reserved'}


ARITH_FAST, ARITH_EXACT = 0, 1
TAIL_AUTO, TAIL_SINGLE, TAIL_TWO_PHASE, TAIL_CAPS = 0, 1, 2, 3
ENGINE_AUTO, ENGINE_SCALAR, ENGINE_MMA, ENGINE_SINGLE = 0, 1, 2, 3


class SpciesCudaError(RuntimeError):
    pass


def _dptr(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


class CudaSolver:
    def __init__(self, so_path, spec=None):
        self.so_path = so_path
        self.spec = spec
        self.lib = ctypes.CDLL(so_path)
        L = self.lib
        L.spcies_cuda_solver_name.restype = c_char_p
        L.spcies_cuda_save_name.restype = c_char_p
        L.spcies_cuda_precision.restype = c_char_p
        L.spcies_cuda_last_error.restype = c_char_p
        L.spcies_cuda_sol_doubles.restype = c_long
        self.solver_name = L.spcies_cuda_solver_name().decode()
        self.save_name = L.spcies_cuda_save_name().decode()
        self.precision = L.spcies_cuda_precision().decode()
        L.spcies_cuda_arithmetic.restype = c_char_p
        self.arithmetic = L.spcies_cuda_arithmetic().decode()
        nn, mm, NN = c_int(), c_int(), c_int()
        L.spcies_cuda_dims(ctypes.byref(nn), ctypes.byref(mm), ctypes.byref(NN))
        self.n, self.m, self.N = nn.value, mm.value, NN.value
        self.sol_doubles = int(L.spcies_cuda_sol_doubles())
        self.func_name = spec.func_name if spec is not None else self._guess_func()
        self.has_r = hasattr(L, 'ellipMPC_ADMM_soc')
        # three references (x_re, x_rs, x_rc) / (u_re, u_rs, u_rc): ellipHMPC, header_ellipHMPC_ADMM_C.h:24
        self.nref = 3 if (spec is not None and 'xrs' in spec.extra_inputs) else 1
        # TIME_VARYING solvers take the model per call / per instance: A, B (column-major), Q, R, LB, UB (code_laxMPC_FISTA_C.c:19)
        self.tv = spec is not None and 'A_in' in spec.extra_inputs
        self._single = getattr(L, self.func_name)
        self._single.restype = None
        self._batch = getattr(L, self.func_name + '_batch')
        self._batch.restype = c_int
        self._closed_loop = getattr(L, self.func_name + '_closed_loop', None)
        if self._closed_loop is not None:
            self._closed_loop.restype = c_int
        self.sol_fields = tuple(spec.sol_fields) if spec is not None else ()

    def _guess_func(self):
        for name in ('laxMPC_FISTA', 'laxMPC_ADMM', 'equMPC_FISTA', 'equMPC_ADMM', 'ellipMPC_ADMM_soc',
                     'ellipMPC_ADMM', 'MPCT_EADMM', 'HMPC_ADMM'):
            if hasattr(self.lib, name + '_batch'):
                return name
        raise SpciesCudaError('no known solver symbol in ' + self.so_path)

    # ------------------------------------------------------------------------------------------
    def device_count(self):
        return int(self.lib.spcies_cuda_device_count())

    def last_error(self):
        return self.lib.spcies_cuda_last_error().decode()

    def free(self):
        self.lib.spcies_cuda_free()

    def kernel_attributes(self, arith=ARITH_FAST):
        vals = [c_int() for _ in range(5)]
        rc = self.lib.spcies_cuda_kernel_attributes(int(arith), *[ctypes.byref(v) for v in vals])
        if rc != 0:
            raise SpciesCudaError(f'spcies_cuda_kernel_attributes failed ({rc}): {self.last_error()}')
        keys = ('regs', 'smem_static', 'smem_dynamic', 'block_threads', 'local_bytes')
        return {k: v.value for k, v in zip(keys, vals)}

    def _split_sol(self, raw):
        out, off = {}, 0
        for name, length in self.sol_fields:
            out[name] = raw[..., off:off + length]
            off += length
        for i, name in enumerate(('update_time', 'solve_time', 'polish_time', 'run_time')):
            out[name] = raw[..., off + i]
        return out

    # ------------------------------------------------------------------------------------------
    def _tv_args(self, tv, LB, UB, B):
        """(A, Bm, Q, R) row-major per instance -> the column-major arrays of the reference signature, + LB / UB."""
        if tv is None or LB is None or UB is None:
            raise ValueError('a TIME_VARYING solver needs tv = (A, B, Q, R) and LB, UB')
        A_, B_, Q_, R_ = tv
        Ac = np.ascontiguousarray(np.transpose(np.asarray(A_, dtype=np.float64).reshape(B, self.n, self.n), (0, 2, 1)))
        Bc = np.ascontiguousarray(np.transpose(np.asarray(B_, dtype=np.float64).reshape(B, self.n, self.m), (0, 2, 1)))
        Qc = np.ascontiguousarray(np.asarray(Q_, dtype=np.float64).reshape(B, self.n))
        Rc = np.ascontiguousarray(np.asarray(R_, dtype=np.float64).reshape(B, self.m))
        LBc = np.ascontiguousarray(np.asarray(LB, dtype=np.float64).reshape(B, self.n + self.m))
        UBc = np.ascontiguousarray(np.asarray(UB, dtype=np.float64).reshape(B, self.n + self.m))
        return [Ac, Bc, Qc, Rc, LBc, UBc]

    def solve(self, x0, xr, ur, r=None, tv=None, LB=None, UB=None):
        """Single instance through the reference's own symbol and signature."""
        x0 = np.ascontiguousarray(x0, dtype=np.float64).ravel().copy()
        xrs = [np.ascontiguousarray(a, dtype=np.float64).ravel().copy() for a in (xr if self.nref == 3 else [xr])]
        urs = [np.ascontiguousarray(a, dtype=np.float64).ravel().copy() for a in (ur if self.nref == 3 else [ur])]
        if x0.size != self.n or len(xrs) != self.nref or len(urs) != self.nref or any(a.size != self.n for a in xrs) or \
                any(a.size != self.m for a in urs):
            raise ValueError('x0 / xr / ur have the wrong dimensions')       # Spcies:<F>:nrhs / size checks of the MEX layer
        u = np.zeros(self.m)
        k, e = c_int(0), c_int(0)
        sol = np.zeros(self.sol_doubles)
        args = [_dptr(x0)] + [_dptr(a) for a in xrs] + [_dptr(a) for a in urs]
        if self.tv:
            tva = self._tv_args(tv, LB, UB, 1)
            args += [_dptr(a) for a in tva]
        if self.has_r:
            if r is None:
                raise ValueError('this solver takes the size of the terminal ellipsoid: r is required')
            rr = np.array([float(r)])
            args.append(_dptr(rr))
        args += [_dptr(u), ctypes.byref(k), ctypes.byref(e), _dptr(sol)]
        self._single(*args)
        if e.value == -100:
            raise SpciesCudaError('device failure: ' + self.last_error())
        return u, k.value, e.value, self._split_sol(sol)

    def solve_batch(self, x0, xr, ur, r=None, LB=None, UB=None, arith=ARITH_FAST, device=0, n_devices=1,
                    want_sol=False, out=None, block_threads=0, grid_blocks=0, tail_mode=0, tail_grace=0, engine=0, tail_caps=(),
                    tv=None):
        """B instances through ``<func>_batch``.  Arrays are ``[B, n]`` / ``[B, m]`` (instance-major)."""
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        xrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (xr if self.nref == 3 else [xr])]
        urs = [np.ascontiguousarray(a, dtype=np.float64) for a in (ur if self.nref == 3 else [ur])]
        B = x0.shape[0] if x0.ndim == 2 else 0
        if x0.shape != (B, self.n) or len(xrs) != self.nref or len(urs) != self.nref or \
                any(a.shape != (B, self.n) for a in xrs) or any(a.shape != (B, self.m) for a in urs):
            raise ValueError('x0 / xr / ur must be [B, nn_], [B, nn_], [B, mm_]')
        if self.has_r and r is None:
            raise ValueError('this solver takes the size of the terminal ellipsoid: r [B] is required')
        if out is None:
            u = np.empty((B, self.m))
            k = np.empty(B, dtype=np.int32)
            e = np.empty(B, dtype=np.int32)
        else:
            u, k, e = out
            for a, dt, shp in ((u, np.float64, (B, self.m)), (k, np.int32, (B,)), (e, np.int32, (B,))):
                if not isinstance(a, np.ndarray) or a.dtype != dt or a.shape != shp or not a.flags['C_CONTIGUOUS']:
                    raise ValueError('out = (u float64 [B, mm_], k int32 [B], e_flag int32 [B]), C-contiguous')
        sol = np.zeros((B, self.sol_doubles)) if want_sol else None
        opts = BatchOpts()
        opts.device, opts.n_devices, opts.arith = int(device), int(n_devices), int(arith)
        opts.block_threads, opts.grid_blocks = int(block_threads), int(grid_blocks)
        opts.tail_mode, opts.tail_grace = int(tail_mode), int(tail_grace)
        opts.engine = int(engine)
        for i, c in enumerate(tuple(tail_caps)[:3]):
            opts.tail_caps[i] = int(c)
        keep = []
        tva = None
        if self.tv:
            tva = self._tv_args(tv, LB, UB, B)
            LB = UB = None                    # they travel in the argument list of a TIME_VARYING solver
        if LB is not None or UB is not None:
            LB = np.ascontiguousarray(LB, dtype=np.float64)
            UB = np.ascontiguousarray(UB, dtype=np.float64)
            if LB.shape != (B, self.n + self.m) or UB.shape != LB.shape:
                raise ValueError('LB / UB must be [B, nm_]')
            opts.LB, opts.UB = LB.ctypes.data, UB.ctypes.data
            keep += [LB, UB]
        info = BatchInfo()
        args = [c_long(B), _dptr(x0)] + [_dptr(a) for a in xrs] + [_dptr(a) for a in urs]
        if tva is not None:
            args += [_dptr(a) for a in tva]
        if self.has_r:
            rr = np.ascontiguousarray(np.broadcast_to(np.asarray(r, dtype=np.float64).ravel(), (B,)))
            keep.append(rr)
            args.append(_dptr(rr))
        args += [_dptr(u), _dptr(k), _dptr(e), _dptr(sol), ctypes.byref(opts), ctypes.byref(info)]
        rc = self._batch(*args)
        if rc != 0:
            raise SpciesCudaError(f'{self.func_name}_batch failed ({rc}): {self.last_error()}')
        if want_sol:
            return u, k, e, info.as_dict(), self._split_sol(sol)
        return u, k, e, info.as_dict()

    def closed_loop(self, x0, xr, ur, steps, r=None, warm_start=False, plant_AB=None, arith=ARITH_FAST, device=0, n_devices=1,
                    engine=0, want_x=True):
        """``steps`` sampling times of the closed loop ``u_t = MPC(x_t)``, ``x_{t+1} = A x_t + B u_t`` for every instance (the loop
        of examples/cl_in_C/main_cl_in_C.c:100-117, batched and device-resident) through ``<func>_closed_loop``.
        Returns ``x [steps+1, B, n] (or None), u [steps, B, m], k [steps, B], e_flag [steps, B], info``."""
        if self._closed_loop is None:
            raise SpciesCudaError('this solver has no closed-loop entry point')
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        xr = np.ascontiguousarray(xr, dtype=np.float64)
        ur = np.ascontiguousarray(ur, dtype=np.float64)
        B = x0.shape[0] if x0.ndim == 2 else 0
        if x0.shape != (B, self.n) or xr.shape != (B, self.n) or ur.shape != (B, self.m):
            raise ValueError('x0 / xr / ur must be [B, nn_], [B, nn_], [B, mm_]')
        steps = int(steps)
        x = np.empty((steps + 1, B, self.n)) if want_x else None
        u = np.empty((steps, B, self.m))
        k = np.empty((steps, B), dtype=np.int32)
        e = np.empty((steps, B), dtype=np.int32)
        opts = BatchOpts()
        opts.device, opts.n_devices, opts.arith, opts.engine = int(device), int(n_devices), int(arith), int(engine)
        opts.warm_start = int(warm_start)            # 0 cold, 1 previous dual point, 2 previous dual point shifted by one stage
        keep = []
        if plant_AB is not None:
            pab = np.ascontiguousarray(plant_AB, dtype=np.float64)
            if pab.shape != (self.n, self.n + self.m):
                raise ValueError('plant_AB must be [nn_, nm_]')
            opts.plant_AB = pab.ctypes.data
            keep.append(pab)
        info = BatchInfo()
        args = [c_long(B), c_int(steps), _dptr(x0), _dptr(xr), _dptr(ur)]
        if self.has_r:
            if r is None:
                raise ValueError('this solver takes the size of the terminal ellipsoid: r [B] is required')
            rr = np.ascontiguousarray(np.broadcast_to(np.asarray(r, dtype=np.float64).ravel(), (B,)))
            keep.append(rr)
            args.append(_dptr(rr))
        args += [_dptr(x), _dptr(u), _dptr(k), _dptr(e), ctypes.byref(opts), ctypes.byref(info)]
        rc = self._closed_loop(*args)
        if rc != 0:
            raise SpciesCudaError(f'{self.func_name}_closed_loop failed ({rc}): {self.last_error()}')
        return x, u, k, e, info.as_dict()

    def solve_batch_device(self, B, d_x0, d_xr, d_ur, d_u, d_k, d_e, d_r=None, d_LB=None, d_UB=None, arith=ARITH_FAST,
                           device=0, stream=None, block_threads=0, grid_blocks=0, tail_mode=0, tail_grace=0, engine=0, tail_caps=()):
        """Same call with DEVICE pointers (integers), e.g. ``tensor.data_ptr()``: no copies, kernel only."""
        opts = BatchOpts()
        opts.device, opts.n_devices, opts.arith, opts.device_pointers = int(device), 1, int(arith), 1
        opts.block_threads, opts.grid_blocks = int(block_threads), int(grid_blocks)
        opts.tail_mode, opts.tail_grace = int(tail_mode), int(tail_grace)
        opts.engine = int(engine)
        for i, c in enumerate(tuple(tail_caps)[:3]):
            opts.tail_caps[i] = int(c)
        opts.LB = d_LB
        opts.UB = d_UB
        opts.stream = stream
        info = BatchInfo()
        args = [c_long(B), c_void_p(d_x0), c_void_p(d_xr), c_void_p(d_ur)]
        if self.has_r:
            args.append(c_void_p(d_r))
        args += [c_void_p(d_u), c_void_p(d_k), c_void_p(d_e), None, ctypes.byref(opts), ctypes.byref(info)]
        rc = self._batch(*args)
        if rc != 0:
            raise SpciesCudaError(f'{self.func_name}_batch failed ({rc}): {self.last_error()}')
        return info.as_dict()
