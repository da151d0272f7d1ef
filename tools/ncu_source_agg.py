"""Aggregate warp-stall samples of a kernel by SASS opcode (from an .ncu-rep source page).
    python tools/ncu_source_agg.py rep.ncu-rep regex:kernel"""
import csv
import subprocess
import sys
from collections import defaultdict


def main(path, kernel):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--kernel-name', kernel],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    data = []
    for r in rows:
        if r and r[0] == 'Kernel Name':
            if hdr is not None:
                break
            continue
        if hdr is None:
            hdr = r
            continue
        if len(r) == len(hdr):
            data.append(r)
    ci = {h: i for i, h in enumerate(hdr)}
    agg = defaultdict(lambda: [0, 0, 0])
    for r in data:
        src = r[ci['Source']].strip()
        toks = src.split()
        op = toks[0]
        if op.startswith('@') and len(toks) > 1:
            op = toks[1]
        op = op.split('.')[0]
        a = agg[op]
        a[0] += int(r[ci['# Samples']])
        a[1] += int(r[ci['Instructions Executed']])
        a[2] += 1
    tot = sum(v[0] for v in agg.values())
    toti = sum(v[1] for v in agg.values())
    print('total samples', tot, 'warp instructions', toti)
    for op, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
        print('%-10s samples %7d (%5.1f%%)  executed %11d (%5.1f%%)  static %5d' %
              (op, v[0], 100.0 * v[0] / tot, v[1], 100.0 * v[1] / toti, v[2]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
