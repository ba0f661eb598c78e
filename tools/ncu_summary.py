#!/usr/bin/env python
"""Print the handful of ncu metrics the design notes quote, for every kernel in a .ncu-rep (development tool)."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.avg', 'smsp__average_warp_latency_per_inst_issued.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum', 'sm__inst_executed_pipe_fp64.sum',
        'sm__inst_executed_pipe_lsu.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        # tensor pipe (DMMA.8x8x4 issues here): the counters that cross-check the fraction of the FP64 tensor ceiling
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor.sum', 'sm__inst_executed_pipe_tensor_op_dmma.sum',
        'sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_uniform.sum', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct']


def main(path, grep=None):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('=====', d['Kernel Name'][:70], d.get('Block Size'), d.get('Grid Size'))
        for k in KEYS:
            if d.get(k) not in (None, ''):
                print('   %-80s %s' % (k, d[k]))
        for k in hdr:
            if 'tensor' in k and 'ops_path' not in k and 'attribute' not in k and k not in KEYS and d.get(k) not in (None, '', '0', 'n/a'):     # every other tensor-pipe counter the report holds
                print('   %-80s %s' % (k, d[k]))
            if grep and grep in k:
                print('   %-80s %s' % (k, d[k]))
            if 'warps_issue_stalled' in k and k.endswith('.ratio') and 'not_issued' not in k:
                try:
                    v = float(d[k])
                except ValueError:
                    continue
                if v > 0.05:
                    print('      stall %-40s %.3f' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
