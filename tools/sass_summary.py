"""SASS instruction summary of the hot kernels of libpauxy_b200.so (what proves DMMA / TMA / mbarrier
use and the absence of spills):  python tools/sass_summary.py > profiles/rNN/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_fn import functions  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'pauxy_b200', 'csrc', 'libpauxy_b200.so')
HOT = ['taylor3_kernel', 'taylor2_kernel', 'exx_eri_kernel', 'gemm_tma_kernel', 'theta_kernel', 'cholqr_kernel',
       'exchange_kernel', 'qr_kernel', 'field_kernel', 'weight_kernel', 'energy_kernel', 'comb_plan_kernel',
       'pull_pairs_kernel']
COLS = ['DMMA', 'UBLKCP', 'SYNCS', 'LDGSTS', 'LDS', 'STS', 'LDG', 'STG', 'BAR', 'SHFL', 'REDUX', 'LDL', 'STL']


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout
    return out.splitlines()


def main():
    res = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r'\s*Function (\S+):', line)
        if m:
            cur = m.group(1)
        m = re.search(r'REG:(\d+) STACK:(\d+) SHARED:(\d+)', line)
        if m and cur:
            regs[cur] = tuple(int(x) for x in m.groups())
    rows = []
    for name, body in functions(LIB):
        if not any(h in name for h in HOT):
            continue
        ops = collections.Counter()
        for line in body:
            m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
            if m:
                op = m.group(1)
                ops['CREDUX' if op == 'CREDUX' else op] += 1
        ops['REDUX'] += ops.pop('CREDUX', 0)
        rows.append((name, len(body), ops))
    names = demangle([r[0] for r in rows])
    print('# cuobjdump -sass / -res-usage of pauxy_b200/csrc/libpauxy_b200.so (sm_100a), hot kernels only')
    print('# DMMA = FP64 tensor op (mma.sync.m8n8k4.f64), UBLKCP = TMA bulk copy (cp.async.bulk), SYNCS = mbarrier,')
    print('# LDGSTS = cp.async, REDUX = warp reduction, LDL/STL = local memory (spills), STACK in bytes')
    print('%-88s %6s %4s %5s  %s' % ('kernel', 'instr', 'regs', 'stack', ' '.join('%6s' % c for c in COLS)))
    for (name, n, ops), dn in sorted(zip(rows, names), key=lambda x: x[1]):
        r = regs.get(name, (0, 0, 0))
        short = re.sub(r'\(.*$', '', dn).replace('void ', '').replace('pxb::', '')
        print('%-88s %6d %4d %5d  %s' % (short[:88], n, r[0], r[1], ' '.join('%6d' % ops.get(c, 0) for c in COLS)))


if __name__ == '__main__':
    main()
