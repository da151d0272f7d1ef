"""Top stalled SASS instructions of a kernel from an .ncu-rep source page.
    python tools/ncu_source_top.py rep.ncu-rep regex:taylor [N]"""
import csv
import subprocess
import sys


def main(path, kernel, top=40):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--kernel-name', kernel],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # sections: "Kernel Name" line, header line, data lines
    sect, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'hdr': None, 'data': []}
            sect.append(cur)
        elif cur is not None and cur['hdr'] is None:
            cur['hdr'] = r
        elif cur is not None:
            cur['data'].append(r)
    for s in sect[:1]:
        ci = {h: i for i, h in enumerate(s['hdr'])}
        data = [r for r in s['data'] if len(r) == len(s['hdr'])]
        tot = sum(int(r[ci['# Samples']]) for r in data)
        print(s['name'][:100], 'instructions', len(data), 'samples', tot)
        idx = sorted(range(len(data)), key=lambda k: -int(data[k][ci['# Samples']]))[:top]
        for k in sorted(idx):
            r = data[k]
            print('%5d %-90s %6s %9s' % (k, r[ci['Source']].strip()[:90], r[ci['# Samples']],
                                         r[ci['Instructions Executed']]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
