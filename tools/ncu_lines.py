#!/usr/bin/env python
"""Attribute executed SASS instructions / stall samples of an ncu report to CUDA source lines.

    python tools/ncu_lines.py <report.ncu-rep> <object.o> <mangled kernel name> [warp_stages]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, obj, kname = sys.argv[1:4]
    ws = float(sys.argv[4]) if len(sys.argv) > 4 else 81920000.0
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    sass = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.split('\n')
    start = [i for i, l in enumerate(sass) if l.startswith('.text.' + kname + ':')][0]
    cur, addr2line = None, {}
    for l in sass[start + 1:]:
        if l.startswith('.text.') or l.startswith('//--------------------- .text'):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            addr2line[int(m.group(1), 16)] = cur
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[1]
    iA, iS, iE, iN = h.index('Address'), h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith('stall_')]
    base = None
    byline, samp = collections.Counter(), collections.Counter()
    ops, stalls = collections.defaultdict(collections.Counter), collections.defaultdict(collections.Counter)
    for r in rows[2:]:
        try:
            a, n, s = int(r[iA], 16), int(r[iE]), int(r[iN])
        except (ValueError, IndexError):
            continue
        if base is None:
            base = a
        ln = addr2line.get(a - base)
        byline[ln] += n / ws
        samp[ln] += s
        toks = r[iS].split()
        op = toks[1] if toks[0].startswith('@') else toks[0]
        ops[ln][op.split('.')[0]] += n / ws
        for i, c in stall_cols:
            try:
                stalls[ln][c[6:]] += int(r[i])
            except ValueError:
                pass
    tot = sum(samp.values())
    print('total inst/warp-stage %.1f, samples %d' % (sum(byline.values()), tot))
    for ln, s in samp.most_common(40):
        top = ', '.join('%s %.1f' % kv for kv in ops[ln].most_common(4))
        st = ', '.join('%s %.0f%%' % (k, 100.0 * v / max(sum(stalls[ln].values()), 1)) for k, v in stalls[ln].most_common(3))
        print('%-26s samples %5.1f%%  inst/ws %6.1f | %s | %s' % (ln, 100.0 * s / tot, byline[ln], top, st))


if __name__ == '__main__':
    main()
