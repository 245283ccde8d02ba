"""Summarise an .ncu-rep: one block of headline metrics per profiled kernel launch.
usage: python profiles/ncu_summary.py report.ncu-rep [more metrics...]"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum.pct_of_peak_sustained_elapsed',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum']


def main():
    rep = sys.argv[1]
    want = WANT + sys.argv[2:]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('==', r[hdr.index('Kernel Name')])
        for w in want:
            if w in hdr:
                print('  %-95s %s %s' % (w, r[hdr.index(w)], units[hdr.index(w)]))
        st = []
        for k, h in enumerate(hdr):
            if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued'):
                try:
                    st.append((float(r[k]), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1.0
        print('  stall samples: ' + ', '.join('%s %.0f%%' % (n, 100 * v / tot) for v, n in sorted(st, reverse=True)[:7]))


if __name__ == '__main__':
    main()
