#!/usr/bin/env python
"""Per CUDA source line: stall samples and warp instructions of an ncu report captured with --import-source on.
    python tools/ncu_lines.py x.ncu-rep [top]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = None
    out = []
    fname = ''
    for r in rows:
        if r and r[0] == 'File Path':
            fname = r[1].split('/')[-1]
        if r and r[0] == 'Line No':
            hdr = r
            ix = {h: i for i, h in enumerate(hdr)}
            stall = [(h, i) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
            continue
        if hdr is None or len(r) < len(hdr) or not r[0]:
            continue
        if len(r) > len(hdr):          # a source line with commas / quotes split into extra fields
            extra = len(r) - len(hdr)
            r = [r[0], ','.join(r[1:2 + extra])] + r[2 + extra:]
        samp = int(float(r[ix['# Samples']] or 0))
        inst = int(float(r[ix['Instructions Executed']] or 0))
        st = sorted(((h[6:], int(float(r[i] or 0))) for h, i in stall), key=lambda kv: -kv[1])[:3]
        out.append((samp, inst, fname, r[0], r[1].strip(), st))
    tot = sum(o[0] for o in out)
    print(f'total samples {tot}')
    for samp, inst, fn, ln, src, st in sorted(out, key=lambda x: -x[0])[:top]:
        print(f'{samp:6d} ({100.0 * samp / max(tot, 1):4.1f}%) {inst:9d}  {fn}:{ln:>4s}  {src[:80]:80s} ' +
              ' '.join(f'{k}={v}' for k, v in st if v))


if __name__ == '__main__':
    main()
