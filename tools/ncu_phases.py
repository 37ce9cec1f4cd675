"""Per-phase dynamic instruction / stall-sample shares of one kernel from an ncu --set full report.

    ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv ; python tools/ncu_phases.py src.csv

Phases are the SASS ranges between BAR.SYNC instructions. Also prints the hottest opcodes by executed count.
"""
import collections
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'hdr': None, 'data': []}
            kernels.append(cur)
        elif cur is not None and cur['hdr'] is None:
            cur['hdr'] = r
        elif cur is not None and len(r) > 10:
            cur['data'].append(r)
    for k in kernels[:1]:
        hdr, data = k['hdr'], k['data']
        iS, iE, iN = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
        tot = sum(int(r[iE]) for r in data)
        tots = sum(int(r[iN]) for r in data)
        print(k['name'])
        print('warp instructions executed', tot, ' stall samples', tots, ' SASS lines', len(data))
        seg = acc = accs = start = 0
        for j, r in enumerate(data):
            acc += int(r[iE]); accs += int(r[iN])
            if 'BAR.SYNC' in r[iS] or j == len(data) - 1:
                print('  phase %d  sass[%d:%d]  inst %10d  %5.1f%%   samples %5.1f%%' % (seg, start, j, acc, 100 * acc / tot, 100 * accs / max(tots, 1)))
                seg += 1; acc = accs = 0; start = j + 1
        ops = collections.Counter()
        for r in data:
            s = r[iS].split()
            op = s[1] if s[0].startswith('@') else s[0]
            ops[op.split('.')[0]] += int(r[iE])
        print('  opcode shares:', ', '.join('%s %.1f%%' % (o, 100 * c / tot) for o, c in ops.most_common(top)))


if __name__ == '__main__':
    main(sys.argv[1])
