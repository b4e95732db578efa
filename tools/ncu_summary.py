#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep (run here, no GPU needed): tools/ncu_summary.py rep [kernel-regex]"""
import csv, io, subprocess, sys, re
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
def main():
    rep = sys.argv[1]; rx = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
    for r in data:
        name = r[ix['Kernel Name']]
        if rx and not rx.search(name): continue
        print('---', name[:70], 'grid', r[ix['launch__grid_size']], 'block', r[ix['launch__block_size']])
        for w in WANT:
            if w in ix: print(f"   {w:70s} {r[ix[w]]:>16s} {units[ix[w]]}")
        v = sorted(((float(r[ix[h]].replace(',', '')), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for h in stall), reverse=True)
        print('   stalls/issue: ' + ', '.join(f"{n}={x:.2f}" for x, n in v[:7]))
main()
