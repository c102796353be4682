#!/usr/bin/env python
"""Static SASS opcode mix of one kernel in an object file (quick look before spending GPU time).

    python tools/sass_mix.py <object.o> <substring of the mangled kernel name>
"""
import collections
import re
import subprocess
import sys


def main():
    obj, key = sys.argv[1:3]
    out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
    cur, mix = None, collections.Counter()
    for line in out.split('\n'):
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            continue
        if cur is None or key not in cur:
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m:
            mix[m.group(1)] += 1
    total = sum(mix.values())
    print('total', total)
    fp64 = sum(v for k, v in mix.items() if k in ('DFMA', 'DMUL', 'DADD', 'DSETP', 'MUFU'))
    print('fp64-pipe', fp64)
    for k, v in mix.most_common(30):
        print('%-8s %5d' % (k, v))


if __name__ == '__main__':
    main()
