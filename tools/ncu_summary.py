#!/usr/bin/env python
"""Condenses an .ncu-rep into the per-launch table kept under profiles/ (read here, on the CPU box).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_xxx.csv
"""
import csv
import io
import subprocess
import sys

METRICS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    # a report, or its `ncu -i X.ncu-rep --page raw --csv` dump made on the GPU box (reports with source exceed the 64 MiB
    # that travel back)
    raw = open(rep).read() if rep.endswith('.csv') else subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    cols = [m for m in METRICS if m in hdr]
    with open(out, 'w', newline='') as fh:
        w = csv.writer(fh)
        w.writerow(['kernel'] + ['%s [%s]' % (m, units[hdr.index(m)]) for m in cols])
        for r in rows[2:]:
            w.writerow([r[hdr.index('Kernel Name')]] + [r[hdr.index(m)] for m in cols])
    print('wrote', out, len(rows) - 2, 'launches')


if __name__ == '__main__':
    main()
