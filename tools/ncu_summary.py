"""Summarise an .ncu-rep (raw page) into a few lines per kernel: duration, DMMA /
tensor pipe utilisation, DRAM bytes, L1/L2 hit rates, registers, top stall reasons.
    python tools/ncu_summary.py gpurun_out/prof_exchange.ncu-rep [> profiles/r01/x.txt]"""
import csv
import subprocess
import sys


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name):
        i = col.get(name)
        return r[i] if i is not None and i < len(r) else ''

    for r in rows[2:]:
        print('kernel:', get(r, 'Kernel Name')[:90], ' grid', get(r, 'Grid Size'), 'block', get(r, 'Block Size'))
        for name in ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second',
                     'launch__registers_per_thread', 'launch__occupancy_limit_registers',
                     'sm__warps_active.avg.pct_of_peak_sustained_active',
                     'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
                     'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
                     'sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
                     'smsp__issue_active.avg.pct_of_peak_sustained_active',
                     'dram__bytes_read.sum', 'dram__bytes_write.sum',
                     'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
                     'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
                     'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
                     'smsp__inst_executed.sum']:
            if name in col:
                print('  %-85s %s %s' % (name, get(r, name), units[col[name]]))
        stalls = []
        for h in hdr:
            if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
                try:
                    stalls.append((float(get(r, h)), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print('  top stalls (warps per issue-active):', ', '.join('%s=%.2f' % (n, v) for v, n in stalls[:6]))


if __name__ == '__main__':
    main(sys.argv[1])
