#!/usr/bin/env python
"""bench.py -- batched MPC QP solves/s on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl reference]

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
}
KERNELS = {'laxMPC_FISTA': 'spcies::fista::fista_mma_kernel', 'equMPC_ADMM': 'spcies::admm::admm_mma_kernel',
           'ellipMPC_ADMM_soc': 'spcies::soc::soc_mma_kernel', 'MPCT_EADMM': 'spcies::eadmm::eadmm_mma_kernel',
           'HMPC_SADMM_split': 'spcies::hmpc::hmpc_mma_kernel', 'HMPC_ADMM_split': 'spcies::hmpc::hmpc_mma_kernel'}


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
    p = os.path.join(ROOT, 'profiles', 'r1_fista_mma_ncu_summary.txt')
    if not os.path.exists(p):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--config', default='C2')
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--batch', type=int, default=0, help='override the per-GPU batch (testing only)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    save_name, B, cpu_sample, desc = WORKLOADS[args.config]
    if args.batch:
        B = args.batch
    W = max(args.warmup, 3)
    K = max(args.steps, 1)

    from spcies_b200 import prebuilt, sysmodel
    spec, cfg = prebuilt.spec_for(save_name)
    dims = spec.dims
    metric = 'batched MPC QP solves/sec'
    config = {'workload': desc, 'solver': spec.options.solver_key(), 'formulation': spec.formulation,
              'method': spec.method, 'N': dims['N'], 'nn': dims['n'], 'mm': dims['m'],
              'batch_per_gpu': B, 'tol': spec.define('tol'), 'k_max': spec.define('k_max'),
              'arith': 'fast (FMA / FP64 MMA)', 'engine': 'auto (DMMA tensor-core kernel, 8 instances per warp)', 'l2': 'inputs larger than L2: 3 distinct batches rotated between steps',
              'seed': 'numpy default_rng(100 + 3*rank + i)'}

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == 'reference':
        if rank != 0:
            return
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
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the solver has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at the first collective: keep stdout for the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
            dist.barrier()
            torch.cuda.synchronize(local_rank)
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    dev = torch.device('cuda', local_rank)
    sol, spec, cfg = prebuilt.get(save_name)
    n, m = sol.n, sol.m

    # three distinct batches (3 x 112 MB of inputs > 126 MB L2), device resident and pinned-host copies
    NB = 3
    host, devb = [], []
    for i in range(NB):
        b = sysmodel.synthetic_batch(cfg['sys'], B, seed=100 + 3 * rank + i, with_r=sol.has_r)
        hb = {k: torch.from_numpy(v).pin_memory() for k, v in b.items()}
        host.append(hb)
        devb.append({k: v.to(dev) for k, v in hb.items()})
    d_u = torch.empty((B, m), dtype=torch.float64, device=dev)
    d_k = torch.empty(B, dtype=torch.int32, device=dev)
    d_e = torch.empty(B, dtype=torch.int32, device=dev)
    h_u = torch.empty((B, m), dtype=torch.float64).pin_memory()
    h_k = torch.empty(B, dtype=torch.int32).pin_memory()
    h_e = torch.empty(B, dtype=torch.int32).pin_memory()
    stream = torch.cuda.current_stream(dev)

    def step_dev(i):
        b = devb[i % NB]
        return sol.solve_batch_device(B, b['x0'].data_ptr(), b['xr'].data_ptr(), b['ur'].data_ptr(),
                                      d_u.data_ptr(), d_k.data_ptr(), d_e.data_ptr(), d_r=b['r'].data_ptr() if sol.has_r else None,
                                      device=local_rank, stream=stream.cuda_stream)

    def step_host(i):
        b = host[i % NB]
        return sol.solve_batch(b['x0'].numpy(), b['xr'].numpy(), b['ur'].numpy(), r=b['r'].numpy() if sol.has_r else None,
                               device=local_rank, out=(h_u.numpy(), h_k.numpy(), h_e.numpy()))[3]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    from spcies_b200.sharding import reduce_scalar

    def max_over_ranks(x):
        return reduce_scalar(x, 'max', dev)

    def sum_over_ranks(x):
        return reduce_scalar(x, 'sum', dev)

    # ---- device-resident throughput
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(W):
        step_dev(i)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    infos = []
    tw0 = time.perf_counter()
    ev0.record(stream)
    for i in range(K):
        infos.append(step_dev(W + i))
    ev1.record(stream)
    barrier()
    sampler.windows.append((tw0, time.perf_counter()))
    t_dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    kernel_ms = float(np.mean([x['kernel_ms'] for x in infos]))
    sum_k = float(np.mean([x['sum_k'] for x in infos]))
    n_nc = float(np.mean([x['n_not_converged'] for x in infos]))
    value = world * B * K / (t_dev_ms * 1e-3)

    # ---- end to end through the host-buffer C-ABI call
    for i in range(2):
        step_host(i)
    barrier()
    t0 = time.perf_counter()
    hinfos = [step_host(W + i) for i in range(K)]
    barrier()
    sampler.windows.append((t0, time.perf_counter()))
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = world * B * K / t_e2e

    # ---- parity gate on a subset (every benchmark run) + CPU baseline, rank 0 at N = 1 only
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = min(B, cpu_sample)                    # ~10-15 s on one core
        b0 = {k: v.numpy() for k, v in host[(W + K - 1) % NB].items()}
        rate1, dt1, ur_, kr, er = cpu_reference_leg(save_name, b0, sample, 1)
        cores = os.cpu_count() or 1
        rate_all, dt_all, _, _, _ = cpu_reference_leg(save_name, b0, sample, cores)
        u, k, e = h_u.numpy()[:sample], h_k.numpy()[:sample], h_e.numpy()[:sample]
        same = (k == kr) & (er == 1)                   # converged instances with the same iteration count
        rel = np.abs(u - ur_) / np.maximum(1.0, np.abs(ur_))
        parity = {'compared': int(sample), 'e_flag_mismatch': int((e != er).sum()),
                  'max_abs_dk': int(np.abs(k - kr).max()), 'n_dk_nonzero': int((k != kr).sum()),
                  'u_opt_max_rel_err_same_k': float(rel[same].max()) if same.any() else None, 'tolerance': 1e-9}
        cpu_baseline = {'value': rate1, 'unit': 'solves/s', 'cores': 1, 'kind': 'reference',
                        'sample': f'first {sample} instances of the last timed batch; instantiated reference template, '
                                  f'gcc -O3, DEBUG/MEASURE_TIME off; {dt1:.1f} s',
                        'all_cores': {'value': rate_all, 'cores': cores, 'seconds': dt_all},
                        'mean_k': float(kr.mean())}

    # ---- single-solve latency through the unchanged single-instance symbol (BASELINE.json metric, second half):
    #      p50 over repeated calls at the reference test point, next to the reference C solver called the same way (ctypes)
    single = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and 'status' in cfg:
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
        c_p50 = None
        exe = os.path.join(ROOT, 'harness', 'main_batch')
        if args.config == 'C2' and os.path.exists(exe):
            try:
                out = subprocess.run([exe, '64', '0'], capture_output=True, text=True, timeout=120).stdout
                c_p50 = float(out.split('p50 latency =')[1].split('us')[0])
            except Exception:
                c_p50 = None
        single = {'p50_us': c_p50 if c_p50 is not None else g50, 'p50_us_ctypes': g50, 'p99_us_ctypes': g99, 'k': int(ks),
                  'cpu_reference_us': c_us,
                  'how': 'single-instance symbol (batch of one: H2D, kernel, D2H): p50 of 200 calls from plain C (harness/main_batch) when '
                         'available, p50 / p99 of 1000 calls through ctypes; reference C solver: mean of %d identical solves, one thread' % nrep}

    sum_k_all = sum_over_ranks(sum_k)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    micro = run_microbench()
    peaks, peak_kind = measured_peaks()
    fma = fma_per_instance(spec.options.solver_key(), dims, sum_k, B, spec)
    achieved_tflops = 2.0 * fma / (kernel_ms * 1e-3) / 1e12
    # ceiling of the FP64 datapath: the larger of the DFMA and DMMA issue rates measured in this run (the solver's tensor-core
    # engine issues DMMA.8x8x4; both instruction kinds share the pipe)
    fp64_peak = 2.0 * max(micro.get('fp64_tfma_per_s', 0.0), micro.get('fp64_dmma_tfma_per_s', 0.0)) if micro and 'fp64_tfma_per_s' in micro else None
    io_bytes = B * (8 * (2 * n + m + (1 if sol.has_r else 0)) + 8 * m + 8)
    roofline = {'bound': 'tensor', 'bound_detail': 'FP64 tensor cores (mma.sync m8n8k4 = SASS DMMA.8x8x4; tcgen05 has no FP64 kind); DMMA '
                                                   'shares the FP64 datapath with DFMA / DADD / DSETP, so this is the FP64 issue ceiling',
                'achieved': achieved_tflops, 'peak': fp64_peak, 'unit': 'TFLOP/s',
                'frac': (achieved_tflops / fp64_peak) if fp64_peak else None,
                'peak_source': 'measured in this run by spcies_b200/csrc/microbench.cu (register-only DMMA.8x8x4 / DFMA, full chip, '
                               'whichever is higher); MEASURED_PEAKS.json has no FP64 figure',
                'kernel': '%s (per step: %d launch(es); algorithmic FMA = SURVEY 8(d) count of the reference algorithm, not the '
                          'padded 8x8x4 MMA slots)' % (KERNELS.get(spec.options.solver_key(), '?'), infos[-1]['launches']),
                'kernel_ms': kernel_ms, 'traffic': ncu_traffic(infos[-1]['launches']) if args.config == 'C2' else None,
                'traffic_source': 'dram__bytes_read.sum + dram__bytes_write.sum over the launches of one step, ncu --set full capture of '
                                  'this command (profiles/r1_fista_mma_ncu_summary.txt); algorithmic batch I/O is roofline.hbm',
                'algorithmic_fma_per_launch': fma, 'sum_k_per_launch': sum_k,
                'hbm': {'bound': 'hbm', 'achieved': io_bytes / (kernel_ms * 1e-3) / 1e9, 'peak': peaks['hbm_gbs'],
                        'unit': 'GB/s', 'frac': io_bytes / (kernel_ms * 1e-3) / 1e9 / peaks['hbm_gbs'],
                        'peak_source': peak_kind + ' (MEASURED_PEAKS.json)', 'algorithmic_bytes_per_launch': io_bytes}}
    line = {'metric': metric, 'value': value, 'unit': 'solves/s', 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': t_dev_ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic', 'config': config,
            'e2e': {'value': e2e_value, 'unit': 'solves/s', 'h2d_bytes_per_step': int(hinfos[-1]['h2d_bytes']),
                    'd2h_bytes_per_step': int(hinfos[-1]['d2h_bytes']), 'ms_per_step': 1e3 * t_e2e / K,
                    'kernel_ms': float(np.mean([x['kernel_ms'] for x in hinfos])),
                    'h2d_ms': float(np.mean([x['h2d_ms'] for x in hinfos])),
                    'd2h_ms': float(np.mean([x['d2h_ms'] for x in hinfos]))},
            'gpu_launches': int(sum(x['launches'] for x in infos)) * world,
            'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu_baseline, 'parity': parity, 'single_solve': single,
            'mean_k': sum_k_all / (world * B), 'n_not_converged_per_batch': n_nc,
            'kernel': {'block_threads': infos[-1]['block_threads'], 'grid_blocks': infos[-1]['grid_blocks'],
                       'smem_bytes': infos[-1]['smem_bytes'], 'regs_per_thread': infos[-1]['regs_per_thread'],
                       'launches_per_step': infos[-1]['launches'],
                       'queue_dry_ms': infos[-1]['drain_us'] / 1e3, 'span_ms': infos[-1]['span_us'] / 1e3,
                       'parked_instances': infos[-1]['parked']},
            'microbench': micro}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
