#!/usr/bin/env python
"""Registers / spills / shared memory per kernel from the `-Xptxas -v` logs written by
`python image_segmentation_lab_b200/_build.py --force --ptxas` (build/*.ptxas.log)."""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    pat = sys.argv[1] if len(sys.argv) > 1 else ''
    rows = []
    for log in sorted(glob.glob(os.path.join(ROOT, 'image_segmentation_lab_b200', 'build', '*.ptxas.log'))):
        txt = open(log).read()
        for m in re.finditer(r"Compiling entry function '([^']+)' for 'sm_100a'\n(.*?)\n(ptxas info\s+: Used [^\n]+)", txt, re.S):
            name, mid, used = m.group(1), m.group(2), m.group(3)
            spill = re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads', mid)
            regs = re.search(r'Used (\d+) registers', used)
            smem = re.search(r'(\d+) bytes smem', used)
            rows.append((name, int(regs.group(1)), int(spill.group(2)) if spill else 0, int(spill.group(3)) if spill else 0,
                         int(smem.group(1)) if smem else 0))
    names = subprocess.run(['c++filt'], input='\n'.join(r[0] for r in rows), capture_output=True, text=True).stdout.split('\n')
    print('%5s %6s %6s %6s  %s' % ('regs', 'st_sp', 'ld_sp', 'smem', 'kernel'))
    for (raw, regs, ss, ls, sm), nm in zip(rows, names):
        nm = nm.replace('b200seg::', '').replace('void ', '')
        nm = re.sub(r'\((?:[A-Za-z]+Params|b200seg_\w+)\)$', '', nm)
        if pat in nm:
            print('%5d %6d %6d %6d  %s' % (regs, ss, ls, sm, nm))


if __name__ == '__main__':
    main()
