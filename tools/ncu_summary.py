"""Summarise an .ncu-rep (raw page) into a small CSV of the metrics the roofline discussion uses."""
import csv
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__shared_mem_per_block_static', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.avg', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_bytes.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 'smsp__inst_executed_op_branch.sum',
        'sm__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_pipe_alu.sum',
        'smsp__inst_executed_pipe_fma.sum', 'smsp__inst_executed_pipe_xu.sum', 'smsp__inst_executed_pipe_uniform.sum',
        'smsp__inst_executed_pipe_cbu.sum', 'smsp__inst_executed_pipe_adu.sum', 'sm__sass_inst_executed_op_shared.sum',
        'sm__sass_inst_executed_op_global.sum', 'sm__sass_inst_executed_op_local.sum']


def main(rep, out=None, idx=0, title=''):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2 + idx]
    lines = [f'# {title}', 'metric,unit,value']
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f'{k},{units[i]},{r[i]}')
    stalls = []
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            try:
                stalls.append((float(r[i].replace(',', '')), h))
            except ValueError:
                pass
    for v, h in sorted(stalls, reverse=True)[:8]:
        lines.append(f'{h},inst,{v:.4f}')
    text = '\n'.join(lines) + '\n'
    if out:
        open(out, 'w').write(text)
    print(text)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, int(sys.argv[3]) if len(sys.argv) > 3 else 0,
         sys.argv[4] if len(sys.argv) > 4 else '')
