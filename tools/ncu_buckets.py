"""Executed instructions and warp-stall samples along the SASS of one kernel, in buckets:
    python tools/ncu_buckets.py rep.ncu-rep <substring of kernel name> [nbuckets]
Shows where a kernel spends its instructions (main loop vs epilogue vs set-up)."""
import csv
import subprocess
import sys


def main(path, needle, nb=30):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'hdr': None, 'data': []}
            secs.append(cur)
        elif cur is not None and cur['hdr'] is None:
            cur['hdr'] = r
        elif cur is not None:
            cur['data'].append(r)
    for s in secs:
        if needle not in s['name']:
            continue
        ci = {h: i for i, h in enumerate(s['hdr'])}
        data = [r for r in s['data'] if len(r) == len(s['hdr']) and r[ci['Instructions Executed']].isdigit()]
        ex = [int(r[ci['Instructions Executed']]) for r in data]
        smp = [int(r[ci['# Samples']]) for r in data]
        print(s['name'][:100], 'static', len(data), 'executed', sum(ex), 'samples', sum(smp))
        n = len(data)
        for b in range(nb):
            lo, hi = b * n // nb, (b + 1) * n // nb
            ops = {}
            for r in data[lo:hi]:
                t = r[ci['Source']].strip().split()
                op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
                ops[op] = ops.get(op, 0) + int(r[ci['# Samples']])
            top = sorted(ops.items(), key=lambda x: -x[1])[:4]
            print('%5d-%5d exec %10d (%4.1f%%) samples %6d (%4.1f%%)  %s' % (
                lo, hi, sum(ex[lo:hi]), 100.0 * sum(ex[lo:hi]) / max(sum(ex), 1), sum(smp[lo:hi]),
                100.0 * sum(smp[lo:hi]) / max(sum(smp), 1), ' '.join('%s:%d' % kv for kv in top)))
        break


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
