#!/usr/bin/env python
"""Per-kernel DRAM bytes / executed warp instructions / ncu duration from an .ncu-rep -> profiles/traffic_rN.json
(read by bench.py for roofline.traffic and the issue roofline).

    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep profiles/traffic_r1.json "source description"
"""
import csv
import io
import json
import re
import subprocess
import sys


def short_name(full):
    """'void b200seg::k<float, (int)5, (bool)1>(P)' -> 'k<float, 5, 1>'"""
    n = re.sub(r'^void\s+', '', full)
    n = re.sub(r'\(.*\)$', '', n) if n.endswith(')') and '<' not in n.split('(')[-1] else n
    depth, cut = 0, len(n)
    for i, ch in enumerate(n):       # cut the parameter list: the first '(' at template depth 0
        if ch == '<':
            depth += 1
        elif ch == '>':
            depth -= 1
        elif ch == '(' and depth == 0:
            cut = i
            break
    n = n[:cut]
    n = n.replace('b200seg::', '').replace('(int)', '').replace('(bool)', '')
    return n.strip()


def main():
    rep, out = sys.argv[1], sys.argv[2]
    src = sys.argv[3] if len(sys.argv) > 3 else rep
    # a report, or its `ncu -i X.ncu-rep --page raw --csv` dump made on the GPU box (reports with source exceed the 64 MiB
    # that travel back)
    raw = open(rep).read() if rep.endswith('.csv') else subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]

    def col(r, name, scale_by_unit=True):
        i = hdr.index(name)
        v = float(r[i]) if r[i] not in ('', 'n/a') else 0.0
        u = units[i].lower()
        if scale_by_unit:
            v *= {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'usecond': 1, 'us': 1, 'nsecond': 1e-3, 'ns': 1e-3,
                  'msecond': 1e3, 'ms': 1e3}.get(u, 1)
        return v

    kernels = {}
    for r in rows[2:]:
        name = short_name(r[hdr.index('Kernel Name')])
        if name in kernels:      # keep the first launch of each kernel
            continue
        rd, wr = col(r, 'dram__bytes_read.sum'), col(r, 'dram__bytes_write.sum')
        kernels[name] = {'dram_bytes_read': rd, 'dram_bytes_write': wr, 'dram_bytes': rd + wr,
                         'ncu_us': col(r, 'gpu__time_duration.sum'), 'warp_inst': col(r, 'smsp__inst_executed.sum', False)}
        # the two units a non-HBM-bound kernel can sit on instead: the L1 / shared-memory data pipe and the issue slots
        for key, metric in (('l1_data_pipe_pct', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'),
                            ('issue_active_pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
                            ('registers', 'launch__registers_per_thread')):
            if metric in hdr:
                kernels[name][key] = col(r, metric, False)
    with open(out, 'w') as fh:
        json.dump({'source': src, 'kernels': kernels}, fh, indent=1)
    print('wrote', out, list(kernels))


if __name__ == '__main__':
    main()
