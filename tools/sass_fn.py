"""Prints the SASS of the kernels whose (mangled) name contains a substring, with a per-kernel
instruction histogram:  python tools/sass_fn.py <lib.so> <substring> [--dump]"""
import collections
import re
import subprocess
import sys


def functions(lib):
    out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
        elif name and re.match(r'\s*/\*[0-9a-f]{4,}\*/', line):
            body.append(line)
    if name:
        yield name, body


def main():
    lib, sub = sys.argv[1], sys.argv[2]
    dump = '--dump' in sys.argv
    for name, body in functions(lib):
        if sub not in name:
            continue
        ops = collections.Counter()
        for line in body:
            m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
            if m:
                ops[m.group(1).split('.')[0]] += 1
        print(name, len(body), 'instructions')
        print('  ', ', '.join('%s %d' % kv for kv in ops.most_common(24)))
        if dump:
            print('\n'.join(body))


if __name__ == '__main__':
    main()
