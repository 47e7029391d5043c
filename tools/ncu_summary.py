"""Summarise an `ncu --set full` capture (read here with `ncu -i ... --page raw --csv`) into profiles/.
usage: python tools/ncu_summary.py <report.ncu-rep> <out.txt> [traffic.json workload trees_per_launch]"""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'gpu__time_duration.sum', 'sm__cycles_elapsed.max',
        'sm__cycles_elapsed.avg.per_second', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'l1tex__m_xbar2l1tex_read_bytes.sum.per_second', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor.sum', 'sm__inst_executed_pipe_uniform.sum', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_utcmma.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.per_cycle_active',
        'smsp__inst_executed.avg.per_cycle_active', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__cycles_active.avg',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum']
lines = []
traffic = None
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    lines.append('=' * 100)
    for k in hdr:
        if k in WANT or 'tensor' in k.lower() and 'pct' in k:
            if d.get(k, '') not in ('', '0', '0.0') and not (k not in WANT and ('.min' in k or '.max' in k or '.sum.' in k)):
                lines.append(f'{k:92s} {d[k]:>22s} {u.get(k, "")}')
    def num(k):
        return float(d[k].replace(',', '')) if d.get(k) else 0.0
    def to_bytes(k):
        v, un = num(k), u.get(k, '')
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(un, 1)
    traffic = to_bytes('dram__bytes_read.sum') + to_bytes('dram__bytes_write.sum')
    lines.append(f'{"dram bytes (read + write) of this launch":92s} {traffic:22.0f} byte')
open(out, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines))
if len(sys.argv) > 5:
    json.dump({'workload': sys.argv[4], 'trees_per_launch': int(sys.argv[5]), 'dram_bytes_per_launch': traffic,
               'source': rep.split('/')[-1]}, open(sys.argv[3], 'w'))
