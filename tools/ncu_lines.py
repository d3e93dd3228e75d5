#!/usr/bin/env python
"""Attribute ncu SASS-level counters (``ncu -i X.ncu-rep --page source --csv``) to source lines using
``nvdisasm -g`` of the cubin extracted from the shipped .so (``cuobjdump -xelf all``).

usage: ncu_lines.py <sass.csv> <nvdisasm -g output> <mangled kernel name substring> [top N]
"""
import csv
import re
import sys
from collections import defaultdict


def parse_disasm(path, kernel):
    """-> list of (file, line) per instruction of the kernel, in address order."""
    out = []
    cur = ('?', 0)
    active = False
    insn = re.compile(r'^\s*/\*([0-9a-f]{4,})\*/\s+\S')
    floc = re.compile(r'//## File "([^"]+)", line (\d+)')
    for ln in open(path, errors='replace'):
        if ln.startswith('//--------------------- .text.'):
            active = kernel in ln
            continue
        if ln.startswith('//--------------------- ') and active and '.text.' not in ln:
            active = False
        if not active:
            continue
        m = floc.search(ln)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        if insn.match(ln):
            out.append(cur)
    return out


def main():
    sass_csv, disasm, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(sass_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr = rows[hdr_i]
    ci = hdr.index('Instructions Executed')
    cs = hdr.index('Warp Stall Sampling (All Samples)')
    body = [r for r in rows[hdr_i + 1:] if len(r) > ci]
    locs = parse_disasm(disasm, kernel)
    print('# %d SASS rows in the report, %d instructions in the disassembly' % (len(body), len(locs)))
    n = min(len(body), len(locs))
    inst = defaultdict(int)
    stall = defaultdict(int)
    tot_i = tot_s = 0
    for k in range(n):
        i = int(float(body[k][ci] or 0))
        s = int(float(body[k][cs] or 0))
        inst[locs[k]] += i
        stall[locs[k]] += s
        tot_i += i
        tot_s += s
    print('# total warp instructions %d, stall samples %d' % (tot_i, tot_s))
    print('# by stall samples')
    for loc, s in sorted(stall.items(), key=lambda x: -x[1])[:top]:
        print('%-22s line %5d  stall %6.2f%%  inst %6.2f%%' % (loc[0], loc[1], 100.0 * s / max(tot_s, 1), 100.0 * inst[loc] / max(tot_i, 1)))
    # by function-ish buckets of 1 file
    print('# by instructions')
    for loc, i in sorted(inst.items(), key=lambda x: -x[1])[:top]:
        print('%-22s line %5d  inst %6.2f%%  stall %6.2f%%' % (loc[0], loc[1], 100.0 * i / max(tot_i, 1), 100.0 * stall[loc] / max(tot_s, 1)))


if __name__ == '__main__':
    main()
