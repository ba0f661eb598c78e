#!/usr/bin/env python
"""bench.py -- batched MPC QP solves/s on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl reference] [--no-others]

A *step* is one pass of the hot path over one batch of synthetic instances (SURVEY.md 8(d) generator).
N = 1 workload: configs[1] of BASELINE.json -- laxMPC FISTA, oscillating masses N=10, B = 1,048,576 random
x0 / references.  N > 1 (torchrun, one rank per GPU): every rank solves its own 1 Mi-instance shard, no
collective on the data path (weak scaling); the time is the max over ranks.

value      whole-job solves/s with inputs resident in HBM (device-pointer entry of the C ABI, kernel +
           its launch overhead), timed with CUDA events on the launching stream.
e2e        same metric through the host-buffer C-ABI call (pinned host arrays; H2D and D2H inside the
           timed region) -- the headline against `--impl reference`.
roofline   dominant kernel (the persistent solver kernel; it is the only kernel of a step):
           algorithmic FP64 flops / launch duration vs the FP64 FMA ceiling measured by
           spcies_b200/csrc/microbench.cu in the same run (MEASURED_PEAKS.json has no FP64 figure;
           its HBM number bounds the batch I/O, reported in roofline.hbm).
cpu_baseline  the instantiated reference C solver (oracle/_ref, gcc -O3) on this box's host cores,
           bounded sample of the same batch.
other_configs  (N = 1) every other BASELINE configuration and the round-2 solvers under the same contract, 2 timed
           steps each: value, e2e, roofline (from that run's sum of k and kernel time), cpu_baseline, parity.
c5_sharded BASELINE.json configs[4] as stated: the N = 50 MPCT EADMM / HMPC batch of 8 Mi instances cut into N contiguous
           shards (strong scaling, one step), per rank and -- for N > 1 -- through ONE C-ABI call with
           spcies_batch_opts.n_devices = N from a single process (the product API's own multi-GPU path).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # config -> (prebuilt solver, per-GPU batch, CPU-baseline sample, description).  C2 is the headline (BASELINE.json configs[1], the
    # default); the others are the remaining BASELINE configurations, measured under the same contract with --config.
    'C2': ('C2_laxMPC_FISTA', 1 << 20, 1 << 19, 'laxMPC FISTA oscillating masses N=10, 1Mi random x0/references per GPU'),
    'C3': ('C3_equMPC_ADMM', 1 << 20, 1 << 14, 'equMPC ADMM oscillating masses N=20, 1Mi-instance batch per GPU, double'),
    'C4': ('C4_ellipMPC_ADMM_soc', 1 << 20, 1 << 14, 'ellipMPC ADMM_soc (proj_SOC terminal constraint) N=10, 1Mi-instance batch per GPU'),
    'C5a': ('C5a_HMPC_SADMM_split', 1 << 17, 512, 'HMPC SADMM_split N=50, 128Ki-instance shard per GPU'),
    'C5b': ('C5b_MPCT_EADMM', 1 << 20, 1 << 14, 'MPCT EADMM N=50, 1Mi-instance shard per GPU'),
    'C3f': ('C3f_equMPC_ADMM', 1 << 20, 1 << 14, 'equMPC ADMM oscillating masses N=20, 1Mi-instance batch per GPU, precision = float '
                                                 '(float constants, double arithmetic: what the reference float solver computes)'),
    'C3ff': ('C3ff_equMPC_ADMM', 1 << 20, 1 << 14, 'equMPC ADMM oscillating masses N=20, 1Mi-instance batch per GPU, precision = float with '
                                                   'true single-precision arithmetic (float_arithmetic)'),
    # solvers added in round 2, at the reference tests' problem (N = 10) with the default tolerances
    'C6': ('C6_MPCT_ADMM_cs', 1 << 18, 1 << 12, 'MPCT ADMM_cs (extended state space) N=10, 256Ki-instance batch per GPU'),
    'C7': ('C7_HMPC_ADMM', 1 << 18, 1 << 12, 'HMPC ADMM (non-split, the toolbox default for HMPC) N=10, 256Ki-instance batch per GPU'),
    'C8': ('C8_MPCT_ADMM_semiband', 1 << 18, 1 << 12, 'MPCT ADMM_semiband (banded + low-rank QP step) N=10, 256Ki-instance batch per GPU'),
}
# configurations measured next to the headline in the default run (2 timed steps each): `other_configs` of the JSON line
OTHER_CONFIGS = ('C3', 'C3f', 'C3ff', 'C4', 'C5a', 'C5b', 'C6', 'C7', 'C8')
KERNELS = {'laxMPC_FISTA': 'spcies::fista::fista_mma_kernel', 'equMPC_ADMM': 'spcies::admm::admm_mma_kernel',
           'ellipMPC_ADMM_soc': 'spcies::soc::soc_mma_kernel', 'MPCT_EADMM': 'spcies::eadmm::eadmm_mma_kernel',
           'HMPC_SADMM_split': 'spcies::dense::dense_mma_kernel<hmpc::Engine>', 'HMPC_ADMM_split': 'spcies::dense::dense_mma_kernel<hmpc::Engine>',
           'MPCT_ADMM_cs': 'spcies::dense::dense_mma_kernel<mpct_cs::Engine>', 'HMPC_ADMM': 'spcies::dense::dense_mma_kernel<hmpc_ns::Engine>',
           'ellipHMPC_ADMM': 'spcies::dense::dense_mma_kernel<hmpc_ns::Engine>',
           'MPCT_ADMM_semiband': 'spcies::dense::dense_mma_kernel<mpct_sb::Engine>'}


def fma_per_instance(solver_name, dims, sum_k, B, spec=None):
    """Algorithmic FMA count of the reference algorithm, not the padded MMA slots."""
    n, m, N = dims['n'], dims['m'], dims['N']
    nm = n + m
    F_W = 2 * (N * n * (n - 1) // 2 + (N - 1) * n * n)          # banded-Cholesky solve, forward + backward
    if solver_name == 'laxMPC_FISTA':                            # FMA(k) = (k+1) F_zr + k F_W
        F_zr = 2 * (m * n + (N - 1) * n * nm)
        return (sum_k + B) * F_zr + sum_k * F_W
    if solver_name == 'equMPC_ADMM':
        return sum_k * (F_W + 3 * (N - 1) * n * nm + 2 * m * n)
    if solver_name == 'ellipMPC_ADMM_soc':
        nnz = sum(len(np.ravel(spec.const(c))) for c in ('GhHhi_val', 'Hhi_val', 'HhiGh_val')) + 2 * len(np.ravel(spec.const('L_val')))
        return sum_k * nnz
    if solver_name == 'MPCT_EADMM':
        return sum_k * (F_W + 2 * (N + 1) * n * nm + nm * nm)
    if solver_name in ('HMPC_SADMM_split', 'HMPC_ADMM_split'):
        NP = dims['dim'] + dims['n_s']
        return sum_k * (NP * NP + NP * n)
    if solver_name == 'MPCT_ADMM_cs':                            # the reference's sparse chain: nnz(AHi) + 2 nnz(L) + nnz(Hi) + nnz(HiA)
        nnz = sum(len(np.ravel(spec.const(c))) for c in ('AHi_val', 'Hi_val', 'HiA_val')) + 2 * len(np.ravel(spec.const('L_val')))
        return sum_k * nnz
    if solver_name == 'MPCT_ADMM_semiband':
        # the reference's three Woodbury-structured solves per iteration (code_MPCT_ADMM_semiband_C.c:191-770): 4 block-diagonal
        # products, 2 M_hat + 2 U_hat sparse products, 2 banded Cholesky solves over N + 2 blocks, M_tilde / U_tilde, G and G'
        blk = (N + 1) * (n * n + m * m)
        mhat = 2 * ((N + 1) * (n * n + m * m))
        uhat = N * (n * n + m * m)
        chol = 2 * ((N + 2) * n * (n - 1) // 2 + (N + 1) * n * n)
        mt = 2 * (n + m) * (N + 2) * n
        gg = 2 * (N + 1) * n * (n + m)
        return sum_k * (4 * blk + 2 * mhat + 2 * uhat + 2 * chol + 2 * mt + gg)
    if solver_name in ('HMPC_ADMM', 'ellipHMPC_ADMM'):           # dense M1, M2 + the two sparse products with C
        d = dims['dim']
        return sum_k * (d * d + d * n + len(np.ravel(spec.const('C_val'))) + len(np.ravel(spec.const('Ct_val'))))
    raise KeyError(solver_name)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []          # (host time of arrival, csv line)
        self.windows = []        # [t0, t1] of the timed regions

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ts, ln in self.lines:
            if self.windows and not any(a <= ts <= b + 0.03 for a, b in self.windows):
                continue             # keep only the samples taken while a timed region was running
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nme, val in zip(names, f[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(nme)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'power_w_max': float(max(pw)), 'samples': len(sm)}


def run_microbench():
    exe = os.path.join(ROOT, 'generated_solvers', 'spcies_microbench')
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout.strip().splitlines()[-1]
        return json.loads(out)
    except Exception as ex:                                    # pragma: no cover
        return {'error': str(ex)}


def ncu_traffic(launches_per_step):
    """DRAM bytes (read + write) of one step from the committed `ncu --set full` summary of the same command
    (profiles/r1_fista_mma_ncu_summary.txt, written by tools/ncu_summary.py): sum over the launches of one step."""
    p = None
    for rnd in ('r2', 'r1'):                                     # the latest round's capture
        q = os.path.join(ROOT, 'profiles', rnd + '_fista_mma_ncu_summary.txt')
        if os.path.exists(q):
            p = q
            break
    if p is None:
        return None
    per_kernel, cur = [], None
    for ln in open(p):
        if ln.startswith('====='):
            cur = 0.0
            per_kernel.append(cur)
        elif 'dram__bytes_read.sum' in ln or 'dram__bytes_write.sum' in ln:
            per_kernel[-1] += float(ln.split()[-1]) * 1e6          # ncu prints Mbyte
    if len(per_kernel) < launches_per_step or launches_per_step <= 0:
        return None
    return sum(per_kernel[:launches_per_step])


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p)), 'measured'
    return {'hbm_gbs': 6650.0}, 'fallback'


def cpu_reference_leg(save_name, batch, sample, threads):
    """Time the instantiated reference C solver (oracle/_ref) on a bounded sample.  Checker / baseline only."""
    from oracle import refs
    ref = refs.get(save_name)[0]
    x0, xr, ur = batch['x0'][:sample], batch['xr'][:sample], batch['ur'][:sample]
    kw = {'r': batch['r'][:sample]} if 'r' in batch else {}
    w = min(sample, 64)
    ref.solve_batch(x0[:w], xr[:w], ur[:w], threads=1, **({'r': batch['r'][:w]} if 'r' in batch else {}))   # warm the code path
    t0 = time.perf_counter()
    u, k, e = ref.solve_batch(x0, xr, ur, threads=threads, **kw)
    dt = time.perf_counter() - t0
    return sample / dt, dt, u, k, e


class Ctx:
    """Process / device context of one bench run (one rank per GPU under torchrun)."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get('RANK', '0'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        if not torch.cuda.is_available():
            raise SystemExit('bench.py: no CUDA device -- the solver has no CPU fallback')
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            # NCCL prints its version banner on stdout at the first collective: keep stdout for the one JSON line
            sys.stdout.flush()
            saved_stdout = os.dup(1)
            os.dup2(2, 1)
            try:
                import datetime
                dist.init_process_group('nccl', device_id=torch.device('cuda', self.local_rank), timeout=datetime.timedelta(minutes=30))
                dist.barrier()
                torch.cuda.synchronize(self.local_rank)
                self.host_group = dist.new_group(backend='gloo', timeout=datetime.timedelta(minutes=30))     # host-side waits that leave the GPUs idle (host_barrier)
            finally:
                sys.stdout.flush()
                os.dup2(saved_stdout, 1)
                os.close(saved_stdout)
        self.dev = torch.device('cuda', self.local_rank)
        self.stream = torch.cuda.current_stream(self.dev)
        self.flush_buf = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def host_barrier(self):
        """Barrier on the host only: waiting ranks launch nothing (an NCCL barrier keeps a kernel spinning on every waiting GPU,
        which would time-slice with another process's work on that GPU)."""
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier(group=self.host_group)

    def max_over_ranks(self, x):
        from spcies_b200.sharding import reduce_scalar
        return reduce_scalar(x, 'max', self.dev)

    def sum_over_ranks(self, x):
        from spcies_b200.sharding import reduce_scalar
        return reduce_scalar(x, 'sum', self.dev)

    def flush_l2(self):
        """Write a buffer larger than the 126 MB L2 (between timed steps of the configurations whose inputs are smaller)."""
        if self.flush_buf is None:
            self.flush_buf = self.torch.empty(256 << 20, dtype=self.torch.uint8, device=self.dev)
        self.flush_buf.fill_(1)


def parity_block(spec, u, k, e, ur_, kr, er, tol):
    """u_opt / k / e_flag of the timed run against the reference C solver on the same instances.  `tol` is the north-star gate
    (1e-9 relative in double, 1e-5 in float); relative = |u - v| / max(|v|, floor), element-wise, floor = 1e-3 (double) or the input range 0.8 (float).  Instances that hit k_max
    (e_flag = -1) are reported separately, not masked."""
    floor = 1e-3 if tol < 1e-6 else 0.8      # float: relative to the input range (a float iterate cannot be 1e-5 relative to an entry of 1e-3)
    rel = np.abs(u - ur_) / np.maximum(floor, np.abs(ur_))
    ab = np.abs(u - ur_)
    same = k == kr
    conv = er == 1
    mx = lambda a, msk: float(a[msk].max()) if msk.any() else None
    return {'compared': int(len(k)), 'e_flag_mismatch': int((e != er).sum()),
            'max_abs_dk': int(np.abs(k - kr).max()), 'n_dk_nonzero': int((~same).sum()),
            'u_opt_max_rel_err': mx(rel, same & conv), 'u_opt_max_abs_err': mx(ab, same & conv),
            'n_not_converged': int((~conv).sum()),
            'u_opt_max_rel_err_not_converged': mx(rel, same & ~conv), 'u_opt_max_abs_err_not_converged': mx(ab, same & ~conv),
            'u_opt_max_abs_err_dk_nonzero': mx(ab, ~same),
            'tolerance': tol, 'relative_floor': floor,
            'pass': bool((e == er).all() and np.abs(k - kr).max() <= 1 and (mx(rel, same & conv) or 0.0) <= tol)}


def measure(cx, config_name, W, K, micro, peaks, peak_kind, batch=0, seeds_base=100, cpu=True, flush=False, nb=3, cpu_scale=1.0):
    """One configuration under the bench contract: device-resident value, host-buffer e2e, roofline of its kernel, parity and
    CPU baseline (rank 0 at N = 1).  Returns (dict for the JSON line, extras)."""
    from spcies_b200 import prebuilt, sysmodel
    torch = cx.torch
    save_name, B, cpu_sample, desc = WORKLOADS[config_name]
    if batch:
        B = batch
    sol, spec, cfg = prebuilt.get(save_name)
    dims = spec.dims
    n, m = sol.n, sol.m
    host, devb = [], []
    for i in range(nb):
        b = sysmodel.synthetic_batch(cfg['sys'], B, seed=seeds_base + 3 * cx.rank + i, with_r=sol.has_r)
        hb = {k: torch.from_numpy(v).pin_memory() for k, v in b.items()}
        host.append(hb)
        devb.append({k: v.to(cx.dev) for k, v in hb.items()})
    d_u = torch.empty((B, m), dtype=torch.float64, device=cx.dev)
    d_k = torch.empty(B, dtype=torch.int32, device=cx.dev)
    d_e = torch.empty(B, dtype=torch.int32, device=cx.dev)
    h_u = torch.empty((B, m), dtype=torch.float64).pin_memory()
    h_k = torch.empty(B, dtype=torch.int32).pin_memory()
    h_e = torch.empty(B, dtype=torch.int32).pin_memory()

    def step_dev(i):
        b = devb[i % nb]
        return sol.solve_batch_device(B, b['x0'].data_ptr(), b['xr'].data_ptr(), b['ur'].data_ptr(),
                                      d_u.data_ptr(), d_k.data_ptr(), d_e.data_ptr(), d_r=b['r'].data_ptr() if sol.has_r else None,
                                      device=cx.local_rank, stream=cx.stream.cuda_stream)

    def step_host(i):
        b = host[i % nb]
        return sol.solve_batch(b['x0'].numpy(), b['xr'].numpy(), b['ur'].numpy(), r=b['r'].numpy() if sol.has_r else None,
                               device=cx.local_rank, out=(h_u.numpy(), h_k.numpy(), h_e.numpy()))[3]

    # ---- device-resident throughput
    for i in range(W):
        step_dev(i)
    cx.barrier()
    infos = []
    tw0 = time.perf_counter()
    if not flush:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(cx.stream)
        for i in range(K):
            infos.append(step_dev(W + i))
        ev1.record(cx.stream)
        cx.barrier()
        t_local = ev0.elapsed_time(ev1)
    else:
        evs = []
        for i in range(K):
            cx.flush_l2()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(cx.stream)
            infos.append(step_dev(W + i))
            b_.record(cx.stream)
            evs.append((a, b_))
        cx.barrier()
        t_local = sum(a.elapsed_time(b_) for a, b_ in evs)
    window_dev = (tw0, time.perf_counter())
    t_dev_ms = cx.max_over_ranks(t_local)
    kernel_ms = float(np.mean([x['kernel_ms'] for x in infos]))
    sum_k = float(np.mean([x['sum_k'] for x in infos]))
    n_nc = float(np.mean([x['n_not_converged'] for x in infos]))
    value = cx.world * B * K / (t_dev_ms * 1e-3)

    # ---- end to end through the host-buffer C-ABI call (pinned host arrays, H2D and D2H inside the timed region)
    KH = K if not flush else max(1, K // 2)
    for i in range(2 if not flush else 1):
        step_host(i)
    cx.barrier()
    t0 = time.perf_counter()
    hinfos = [step_host(W + i) for i in range(KH)]
    cx.barrier()
    window_host = (t0, time.perf_counter())
    t_e2e = cx.max_over_ranks(time.perf_counter() - t0)
    e2e_value = cx.world * B * KH / t_e2e

    # ---- parity gate on a subset of the last timed batch + CPU baseline, rank 0 at N = 1 only
    cpu_baseline, parity = None, None
    if cx.rank == 0 and cx.world == 1 and cpu:
        sample = max(64, int(min(B, cpu_sample) * cpu_scale))
        b0 = {k: v.numpy() for k, v in host[(W + KH - 1) % nb].items()}
        rate1, dt1, ur_, kr, er = cpu_reference_leg(save_name, b0, sample, 1)
        cores = os.cpu_count() or 1
        rate_all, dt_all, _, _, _ = cpu_reference_leg(save_name, b0, sample, cores)
        u, k, e = h_u.numpy()[:sample], h_k.numpy()[:sample], h_e.numpy()[:sample]
        parity = parity_block(spec, u, k, e, ur_, kr, er, 1e-5 if sol.arithmetic == 'float' else 1e-9)
        cpu_baseline = {'value': rate1, 'unit': 'solves/s', 'cores': 1, 'kind': 'reference',
                        'sample': f'first {sample} instances of the last timed batch; instantiated reference template, '
                                  f'gcc -O3, DEBUG/MEASURE_TIME off; {dt1:.1f} s',
                        'all_cores': {'value': rate_all, 'cores': cores, 'seconds': dt_all},
                        'mean_k': float(kr.mean())}

    sum_k_all = cx.sum_over_ranks(sum_k)
    fma = fma_per_instance(spec.options.solver_key(), dims, sum_k, B, spec)
    achieved_tflops = 2.0 * fma / (kernel_ms * 1e-3) / 1e12
    is_float = sol.arithmetic == 'float'
    if micro and 'fp64_tfma_per_s' in micro:
        fp_peak = 2.0 * (micro.get('fp32_tfma_per_s', 0.0) if is_float else
                         max(micro.get('fp64_tfma_per_s', 0.0), micro.get('fp64_dmma_tfma_per_s', 0.0)))
    else:
        fp_peak = None
    io_bytes = B * (8 * (2 * n + m + (1 if sol.has_r else 0)) + 8 * m + 8)
    kname = KERNELS.get(spec.options.solver_key(), '?') if not is_float else 'spcies::persistent_kernel<admm::Solver> (float)'
    roofline = {'bound': 'tensor', 'bound_detail': 'FP64 tensor cores (mma.sync m8n8k4 = SASS DMMA.8x8x4; tcgen05 has no FP64 kind); DMMA '
                                                   'shares the FP64 datapath with DFMA / DADD / DSETP, so this is the FP64 issue ceiling'
                                                   if not is_float else 'FP32 FMA issue (one-thread-per-instance float kernel)',
                'achieved': achieved_tflops, 'peak': fp_peak, 'unit': 'TFLOP/s',
                'frac': (achieved_tflops / fp_peak) if fp_peak else None,
                'peak_source': 'measured in this run by spcies_b200/csrc/microbench.cu (register-only DMMA.8x8x4 / DFMA / FFMA, full chip); '
                               'MEASURED_PEAKS.json has no FP64 / FP32 FMA figure',
                'kernel': '%s (per step: %d launch(es); algorithmic FMA = SURVEY 8(d) count of the reference algorithm, not the '
                          'padded 8x8x4 MMA slots)' % (kname, infos[-1]['launches']),
                'kernel_ms': kernel_ms, 'traffic': None,
                'algorithmic_fma_per_launch': fma, 'sum_k_per_launch': sum_k,
                'hbm': {'bound': 'hbm', 'achieved': io_bytes / (kernel_ms * 1e-3) / 1e9, 'peak': peaks['hbm_gbs'],
                        'unit': 'GB/s', 'frac': io_bytes / (kernel_ms * 1e-3) / 1e9 / peaks['hbm_gbs'],
                        'peak_source': peak_kind + ' (MEASURED_PEAKS.json)', 'algorithmic_bytes_per_launch': io_bytes}}
    out = {'value': value, 'unit': 'solves/s', 'steps': K, 'warmup': W, 'ms_per_step': t_dev_ms / K,
           'config': {'workload': desc, 'solver': spec.options.solver_key(), 'N': dims['N'], 'batch_per_gpu': B,
                      'tol': spec.define('tol', spec.define('tol_p')), 'k_max': spec.define('k_max'), 'precision': sol.precision,
                      'arithmetic': sol.arithmetic,
                      'l2': ('L2 flushed (256 MiB write) before every timed step' if flush else
                             'inputs larger than L2: %d distinct batches rotated between steps' % nb)},
           'dtype': 'f32' if is_float else 'f64',
           'e2e': {'value': e2e_value, 'unit': 'solves/s', 'h2d_bytes_per_step': int(hinfos[-1]['h2d_bytes']),
                   'd2h_bytes_per_step': int(hinfos[-1]['d2h_bytes']), 'ms_per_step': 1e3 * t_e2e / KH, 'steps': KH,
                   'kernel_ms': float(np.mean([x['kernel_ms'] for x in hinfos])),
                   'h2d_ms': float(np.mean([x['h2d_ms'] for x in hinfos])),
                   'd2h_ms': float(np.mean([x['d2h_ms'] for x in hinfos]))},
           'gpu_launches': int(sum(x['launches'] for x in infos)) * cx.world,
           'roofline': roofline, 'cpu_baseline': cpu_baseline, 'parity': parity,
           'mean_k': sum_k_all / (cx.world * B), 'n_not_converged_per_batch': n_nc,
           'kernel': {'block_threads': infos[-1]['block_threads'], 'grid_blocks': infos[-1]['grid_blocks'],
                      'smem_bytes': infos[-1]['smem_bytes'], 'regs_per_thread': infos[-1]['regs_per_thread'],
                      'launches_per_step': infos[-1]['launches'],
                      'queue_dry_ms': infos[-1]['drain_us'] / 1e3, 'span_ms': infos[-1]['span_us'] / 1e3,
                      'parked_instances': infos[-1]['parked']}}
    extras = dict(sol=sol, spec=spec, cfg=cfg, windows=[window_dev, window_host], infos=infos)
    del devb, host
    return out, extras


C5_TOTAL = 1 << 23          # BASELINE.json configs[4]: 8M-instance batch sharded over 1/2/4/8 B200
C5_CHUNK = 1 << 20          # the batch is defined chunk by chunk (seed 200 + chunk index) so that it is the same for every N


def c5_sharded(cx, config_name, total, one_process=True):
    """BASELINE.json configs[4]: the N = 50 batch of `total` instances cut into `world` contiguous shards (strong scaling), one
    step, (i) one rank per GPU, device-resident and through host buffers, and (ii) through the product API's own multi-GPU path:
    ONE host-buffer call with spcies_batch_opts.n_devices = world from rank 0 (one host thread per device inside the library)."""
    from spcies_b200 import prebuilt, sysmodel
    torch = cx.torch
    save_name = WORKLOADS[config_name][0]
    sol, spec, cfg = prebuilt.get(save_name)
    n, m = sol.n, sol.m
    nchunk = max(1, total // C5_CHUNK)
    chunk = total // nchunk
    per = nchunk // cx.world if nchunk >= cx.world else 0
    if per == 0:
        return {'skipped': 'fewer chunks than ranks'}
    mine = range(cx.rank * per, (cx.rank + 1) * per)
    parts = [sysmodel.synthetic_batch(cfg['sys'], chunk, seed=200 + c) for c in mine]
    hb = {k: torch.from_numpy(np.concatenate([p[k] for p in parts])).pin_memory() for k in ('x0', 'xr', 'ur')}
    B = per * chunk
    db = {k: v.to(cx.dev) for k, v in hb.items()}
    d_u = torch.empty((B, m), dtype=torch.float64, device=cx.dev)
    d_k = torch.empty(B, dtype=torch.int32, device=cx.dev)
    d_e = torch.empty(B, dtype=torch.int32, device=cx.dev)
    h_u = torch.empty((B, m), dtype=torch.float64).pin_memory()
    h_k = torch.empty(B, dtype=torch.int32).pin_memory()
    h_e = torch.empty(B, dtype=torch.int32).pin_memory()
    warm = min(B, 1 << 14)                              # warm the code path and the device buffers on a small slice
    sol.solve_batch_device(warm, db['x0'].data_ptr(), db['xr'].data_ptr(), db['ur'].data_ptr(), d_u.data_ptr(), d_k.data_ptr(),
                           d_e.data_ptr(), device=cx.local_rank, stream=cx.stream.cuda_stream)
    cx.flush_l2()
    cx.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(cx.stream)
    info = sol.solve_batch_device(B, db['x0'].data_ptr(), db['xr'].data_ptr(), db['ur'].data_ptr(), d_u.data_ptr(), d_k.data_ptr(),
                                  d_e.data_ptr(), device=cx.local_rank, stream=cx.stream.cuda_stream)
    ev1.record(cx.stream)
    cx.barrier()
    t_dev = cx.max_over_ranks(ev0.elapsed_time(ev1)) * 1e-3
    sum_k = cx.sum_over_ranks(info['sum_k'])
    n_nc = cx.sum_over_ranks(info['n_not_converged'])
    cx.barrier()
    t0 = time.perf_counter()
    hinfo = sol.solve_batch(hb['x0'].numpy(), hb['xr'].numpy(), hb['ur'].numpy(), device=cx.local_rank,
                            out=(h_u.numpy(), h_k.numpy(), h_e.numpy()))[3]
    cx.barrier()
    t_e2e = cx.max_over_ranks(time.perf_counter() - t0)
    checksum = cx.sum_over_ranks(float(h_k.numpy().astype(np.int64).sum()))
    out = {'solver': spec.options.solver_key(), 'N': spec.dims['N'], 'total_instances': B * cx.world, 'shards': cx.world,
           'instances_per_shard': B, 'scaling': 'strong', 'steps': 1,
           'value': B * cx.world / t_dev, 'unit': 'solves/s', 'seconds': t_dev,
           'e2e': {'value': B * cx.world / t_e2e, 'unit': 'solves/s', 'seconds': t_e2e,
                   'h2d_bytes_per_step': int(hinfo['h2d_bytes']) * cx.world, 'd2h_bytes_per_step': int(hinfo['d2h_bytes']) * cx.world},
           'mean_k': sum_k / (B * cx.world), 'n_not_converged': n_nc,
           'sum_k_host_call_equals_device_call': bool(abs(checksum - sum_k) < 0.5),
           'data': 'synthetic, chunk c of %d instances seeded default_rng(200 + c): the same batch for every N' % chunk}
    del db
    # ---- the product API's own multi-GPU path: one process, one call, n_devices = world
    if one_process and cx.world > 1:
        res = None
        cx.host_barrier()          # ranks > 0 now wait on the host: rank 0 drives every GPU of the node from one process
        if cx.rank == 0:
            try:
                allp = [sysmodel.synthetic_batch(cfg['sys'], chunk, seed=200 + c) for c in range(per * cx.world)]
                fb = {k: torch.from_numpy(np.concatenate([p[k] for p in allp])).pin_memory() for k in ('x0', 'xr', 'ur')}
                BT = per * cx.world * chunk
                fu = torch.empty((BT, m), dtype=torch.float64).pin_memory()
                fk = torch.empty(BT, dtype=torch.int32).pin_memory()
                fe = torch.empty(BT, dtype=torch.int32).pin_memory()
                w = min(BT, cx.world << 14)
                sol.solve_batch(fb['x0'].numpy()[:w], fb['xr'].numpy()[:w], fb['ur'].numpy()[:w], n_devices=cx.world)
                t0 = time.perf_counter()
                inf = sol.solve_batch(fb['x0'].numpy(), fb['xr'].numpy(), fb['ur'].numpy(), n_devices=cx.world,
                                      out=(fu.numpy(), fk.numpy(), fe.numpy()))[3]
                dt = time.perf_counter() - t0
                res = {'value': BT / dt, 'unit': 'solves/s', 'seconds': dt, 'kernel_ms_max_over_devices': inf['kernel_ms'],
                       'n_devices': inf['n_devices'], 'sum_k': int(inf['sum_k']),
                       'same_sum_k_as_per_rank_run': bool(abs(inf['sum_k'] - sum_k) < 0.5),
                       'how': 'ONE <func>_batch call from one process with spcies_batch_opts.n_devices = %d (host buffers, one host '
                              'thread per device inside the library), wall clock of the call' % cx.world}
            except Exception as ex:                                  # pragma: no cover
                res = {'error': str(ex)}
        cx.host_barrier()
        out['one_process_n_devices'] = res
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--config', default='C2')
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--batch', type=int, default=0, help='override the per-GPU batch (testing only)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-others', action='store_true', help='headline configuration only (skip other_configs / c5_sharded)')
    ap.add_argument('--c5-total', type=int, default=C5_TOTAL)
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    save_name, B, cpu_sample, desc = WORKLOADS[args.config]
    if args.batch:
        B = args.batch
    W = max(args.warmup, 3)
    K = max(args.steps, 1)

    from spcies_b200 import prebuilt
    spec, cfg = prebuilt.spec_for(save_name)
    dims = spec.dims
    metric = 'batched MPC QP solves/sec'
    config = {'workload': desc, 'solver': spec.options.solver_key(), 'formulation': spec.formulation,
              'method': spec.method, 'N': dims['N'], 'nn': dims['n'], 'mm': dims['m'],
              'batch_per_gpu': B, 'tol': spec.define('tol', spec.define('tol_p')), 'k_max': spec.define('k_max'),
              'arith': 'fast (FMA / FP64 MMA)', 'engine': 'auto (DMMA tensor-core kernel, 8 instances per warp)',
              'l2': 'inputs larger than L2: 3 distinct batches rotated between steps',
              'seed': 'numpy default_rng(100 + 3*rank + i)'}

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == 'reference':
        if rank != 0:
            return
        from spcies_b200 import sysmodel
        cores = os.cpu_count() or 1
        sample = min(B, max(256, cpu_sample // 4))
        batch = sysmodel.synthetic_batch(cfg['sys'], sample, seed=100, with_r=bool(spec.extra_inputs))
        rates = []
        for i in range(W + K):
            rate, dt, _, k, e = cpu_reference_leg(save_name, batch, sample, cores)
            if i >= W:
                rates.append((rate, dt))
        total_t = sum(dt for _, dt in rates)
        value = K * sample / total_t
        line = {'metric': metric, 'value': value, 'unit': 'solves/s', 'impl': 'reference', 'n_gpus': args.gpus,
                'steps': K, 'warmup': W, 'ms_per_step': 1e3 * total_t / K, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': config,
                'cpu_baseline': {'value': value, 'unit': 'solves/s', 'cores': cores, 'kind': 'reference',
                                 'sample': f'{sample} instances of the workload per step, gcc -O3 instantiated reference '
                                           f'template, {cores} pthreads over contiguous slices'},
                'e2e': {'value': value, 'unit': 'solves/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'mean_k': float(k.mean()), 'n_not_converged': int((e == -1).sum())}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- native arm (B200)
    cx = Ctx(args)
    micro = run_microbench() if cx.rank == 0 else None
    peaks, peak_kind = measured_peaks()
    sampler = ClockSampler(cx.local_rank)
    if cx.rank == 0:
        sampler.start()
    head, ex = measure(cx, args.config, W, K, micro, peaks, peak_kind, batch=args.batch, cpu=not args.no_cpu_baseline)
    sampler.windows += ex['windows']
    sol, cfg = ex['sol'], ex['cfg']
    if args.config == 'C2':
        head['roofline']['traffic'] = ncu_traffic(ex['infos'][-1]['launches'])
        head['roofline']['traffic_source'] = ('dram__bytes_read.sum + dram__bytes_write.sum over the launches of one step, ncu --set full '
                                              'capture of this command (profiles/*_fista_mma_ncu_summary.txt, latest round); algorithmic '
                                              'batch I/O is roofline.hbm')

    # ---- single-solve latency through the unchanged single-instance symbol (BASELINE.json metric, second half):
    #      p50 over repeated calls at the reference test point, next to the reference C solver called the same way (ctypes)
    single = None
    if cx.rank == 0 and cx.world == 1 and not args.no_cpu_baseline and 'status' in cfg:
        stt = cfg['status']
        rr = cfg['param'].get('r', None) if sol.has_r else None

        def p50_us(fn, reps):
            for _ in range(20):
                fn()
            ts = []
            for _ in range(reps):
                t0 = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t0)
            return float(np.median(ts) * 1e6), float(np.percentile(ts, 99) * 1e6)
        g50, g99 = p50_us(lambda: sol.solve(stt['x'], stt['xr'], stt['ur'], rr), 1000)
        ks = sol.solve(stt['x'], stt['xr'], stt['ur'], rr)[1]
        # the reference C solver at the same point: mean over a single-thread batch of identical instances (no Python per call)
        from oracle import refs as _refs
        ref1 = _refs.get(save_name)[0]
        nrep = 2000 if 'HMPC' not in save_name else 20
        rep = lambda v: np.repeat(np.asarray(v, dtype=np.float64)[None], nrep, axis=0)
        t0 = time.perf_counter()
        ref1.solve_batch(rep(stt['x']), rep(stt['xr']), rep(stt['ur']), threads=1, **({'r': np.full(nrep, rr)} if sol.has_r else {}))
        c_us = (time.perf_counter() - t0) / nrep * 1e6
        # the same symbol called from plain C (harness/main_batch.c, the reference's examples/cl_in_C pattern): no ctypes overhead
        c_p50, c_p50_launch = None, None
        exe = os.path.join(ROOT, 'harness', 'main_batch')
        if args.config == 'C2' and os.path.exists(exe):
            try:
                out = subprocess.run([exe, '64', '0'], capture_output=True, text=True, timeout=120).stdout
                c_p50 = float(out.split('p50 latency =')[1].split('us')[0])
                out = subprocess.run([exe, '64', '0'], capture_output=True, text=True, timeout=120,
                                     env=dict(os.environ, SPCIES_CUDA_SERVER_LINGER_US='0')).stdout
                c_p50_launch = float(out.split('p50 latency =')[1].split('us')[0])
            except Exception:
                pass
        single = {'p50_us': c_p50 if c_p50 is not None else g50, 'p50_us_ctypes': g50, 'p99_us_ctypes': g99, 'k': int(ks),
                  'cpu_reference_us': c_us, 'p50_us_one_launch_per_call': c_p50_launch,
                  'how': 'single-instance symbol: p50 of 200 back-to-back calls from plain C (harness/main_batch) when available, p50 / p99 '
                         'of 1000 calls through ctypes; FISTA solvers: served by the lingering one-CTA kernel through its mapped-memory '
                         'mailbox (SPCIES_CUDA_SERVER_LINGER_US, default 200; 0 = one launch per call), others: a batch of one; '
                         'reference C solver: mean of %d identical solves, one thread' % nrep}

    # ---- closed loop on the device: `steps` sampling times of u = MPC(x), x+ = A x + B u for every instance
    closed = None
    if cx.rank == 0 and cx.world == 1 and args.config == 'C2' and not args.no_others:
        try:
            from spcies_b200 import sysmodel as _sm
            Bc, steps = 1 << 18, 20
            bb = _sm.synthetic_batch(cfg['sys'], Bc, seed=400)
            closed = {'instances': Bc, 'steps': steps, 'solver': spec.options.solver_key()}
            for label, kw in (('cold', dict(warm_start=0)), ('warm_shifted', dict(warm_start=2)), ('warm_unshifted', dict(warm_start=1))):
                sol.closed_loop(bb['x0'][:4096], bb['xr'][:4096], bb['ur'][:4096], 2, want_x=False, **kw)
                dt = 1e30
                for _rep in range(2):          # best of two: the first full-size call also grows the library's device buffers
                    t0 = time.perf_counter()
                    _, _, kk, ee, ci = sol.closed_loop(bb['x0'], bb['xr'], bb['ur'], steps, want_x=False, **kw)
                    dt = min(dt, time.perf_counter() - t0)
                closed[label] = {'mpc_steps_per_s': Bc * steps / dt, 'seconds': dt, 'device_ms': ci['kernel_ms'], 'launches': ci['launches'],
                                 'mean_k': ci['sum_k'] / (Bc * steps), 'n_not_converged': int(ci['n_not_converged'])}
            # the round-1 way: one host-buffer batched call per sampling time, plant step on the host
            AB = np.hstack([cfg['sys']['A'], cfg['sys']['B']])
            x = bb['x0'].copy()
            t0 = time.perf_counter()
            for _ in range(steps):
                uu = sol.solve_batch(x, bb['xr'], bb['ur'])[0]
                x = x @ AB[:, :sol.n].T + uu @ AB[:, sol.n:].T
            dt = time.perf_counter() - t0
            closed['host_loop_of_batched_calls'] = {'mpc_steps_per_s': Bc * steps / dt, 'seconds': dt}
            closed['how'] = ('<func>_closed_loop through the C ABI, host arrays in / trajectories out, wall clock; cold = every sampling time '
                             'starts from lambda = 0 (a loop of reference calls), warm = from the dual point of the previous sampling time '
                             '(unshifted: the `lambda` argument of spcies_laxMPC_FISTA_solver.m:161-164; shifted by one stage)')
        except Exception as exn:                                     # pragma: no cover
            closed = {'error': '%s: %s' % (type(exn).__name__, exn)}

    # ---- the other BASELINE configurations under the same contract, and configs[4] as stated (8 Mi instances, strong scaling)
    others, c5 = None, None
    if not args.no_others and args.config == 'C2':
        if cx.world == 1:
            others = {}
            for name in OTHER_CONFIGS:
                try:
                    o, ex2 = measure(cx, name, 1, 2, micro, peaks, peak_kind, cpu=not args.no_cpu_baseline, flush=True, nb=1,
                                     seeds_base=300, cpu_scale=0.25)
                    sampler.windows += ex2['windows']
                    others[name] = o
                except Exception as exn:                                 # pragma: no cover
                    others[name] = {'error': '%s: %s' % (type(exn).__name__, exn)}
        c5 = {}
        try:
            c5['C5b_MPCT_EADMM'] = c5_sharded(cx, 'C5b', args.c5_total)
            # HMPC N = 50 runs at ~0.1 M solves/s per GPU: the full 8 Mi-instance batch when it is sharded (N >= 2), 1 Mi (1/8 of it) on one GPU
            hm_total = args.c5_total if cx.world >= 2 else min(args.c5_total, 1 << 20)
            c5['C5a_HMPC_SADMM_split'] = c5_sharded(cx, 'C5a', hm_total)
        except Exception as exn:                                         # pragma: no cover
            c5['error'] = '%s: %s' % (type(exn).__name__, exn)

    clocks = sampler.stop() if cx.rank == 0 else None
    if cx.rank != 0:
        if cx.world > 1:
            cx.dist.destroy_process_group()
        return
    head['config'] = dict(config, **head['config'])
    line = {'metric': metric, 'value': head['value'], 'unit': 'solves/s', 'n_gpus': cx.world, 'steps': K, 'warmup': W,
            'ms_per_step': head['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic', 'config': head['config'], 'e2e': head['e2e'], 'gpu_launches': head['gpu_launches'],
            'clocks': clocks, 'roofline': head['roofline'], 'cpu_baseline': head['cpu_baseline'], 'parity': head['parity'],
            'single_solve': single, 'mean_k': head['mean_k'], 'n_not_converged_per_batch': head['n_not_converged_per_batch'],
            'kernel': head['kernel'], 'closed_loop': closed, 'other_configs': others, 'c5_sharded': c5, 'microbench': micro}
    print(json.dumps(line))
    if cx.world > 1:
        cx.dist.destroy_process_group()


if __name__ == '__main__':
    main()
