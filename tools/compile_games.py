#!/usr/bin/env python
"""Compile GDL rulesheets into propnet tables shipped with the package.

Usage:  python tools/compile_games.py [--src DIR] game [game ...]

The rulesheets themselves are NOT copied into this repository: the tables under
``galvanise_zero_b200/games/<game>.npz`` are derived artefacts (gate programs + symbol tables)
produced by ``galvanise_zero_b200.gdl`` from ``<src>/<game>.kif``. Default source directory is the
reference's ``data/rulesheets`` (present in the build container only).
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

from galvanise_zero_b200.gdl.propnet import compile_kif, save_propnet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--src', default='/root/reference/data/rulesheets')
    ap.add_argument('--dst', default=os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                  '..', 'galvanise_zero_b200', 'games'))
    ap.add_argument('games', nargs='+')
    args = ap.parse_args()
    os.makedirs(args.dst, exist_ok=True)
    for g in args.games:
        t0 = time.time()
        pn = compile_kif(os.path.join(args.src, g + '.kif'), verbose=True)
        out = os.path.join(args.dst, g + '.npz')
        save_propnet(pn, out)
        print('%s: %d gates, %d slots -> %s (%.1f KB) in %.1fs' % (
            g, len(pn.gate_out), pn.n_slots, out, os.path.getsize(out) / 1024, time.time() - t0))


if __name__ == '__main__':
    main()
