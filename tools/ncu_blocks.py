"""Basic-block view of one kernel from `ncu --page source --csv`: consecutive SASS instructions with the same executed
count are merged into one block; prints the blocks that carry the most warp instructions, with their opcode mix and stall
samples.   python tools/ncu_blocks.py src.csv [top]"""
import collections
import csv
import sys


def main(path, top=14):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if r and r[0] == 'Address')
    data = [r for r in rows if len(r) > 10 and r[0].startswith('0x')]
    iS, iE, iN = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    tot = sum(int(r[iE]) for r in data)
    tots = sum(int(r[iN]) for r in data)
    blocks, cur = [], None
    for j, r in enumerate(data):
        e = int(r[iE])
        if cur is None or e != cur['e']:
            cur = {'e': e, 'start': j, 'n': 0, 'samples': 0, 'ops': collections.Counter()}
            blocks.append(cur)
        cur['n'] += 1
        cur['samples'] += int(r[iN])
        s = r[iS].split()
        op = s[1] if s[0].startswith('@') else s[0]
        cur['ops'][op.split('.')[0]] += 1
    print('total warp inst', tot, 'samples', tots, 'blocks', len(blocks))
    for b in sorted(blocks, key=lambda b: -b['e'] * b['n'])[:top]:
        print('sass[%5d:%5d] n=%4d exec/instr=%9d  inst %5.1f%%  samples %5.1f%%  %s' % (
            b['start'], b['start'] + b['n'], b['n'], b['e'], 100.0 * b['e'] * b['n'] / tot, 100.0 * b['samples'] / max(tots, 1),
            ' '.join('%s:%d' % kv for kv in b['ops'].most_common(9))))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 14)
