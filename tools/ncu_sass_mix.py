#!/usr/bin/env python
"""Aggregates the SASS view of an ncu report (--set full --import-source on) by opcode: executed warp instructions,
stall samples and the dominant stall reason -- tells which pipe / dependency a kernel waits on.
    python tools/ncu_sass_mix.py x.ncu-rep [top]"""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    by_op = defaultdict(lambda: [0, 0, defaultdict(int)])
    lines = []
    tot_inst = tot_samp = 0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        src = r[ix['Source']].strip()
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?')
        op = op.rstrip(';')
        base = '.'.join(op.split('.')[:2]) if op.startswith(('UTC', 'LDTM', 'STTM', 'MUFU', 'F2FP', 'SYNCS', 'RED', 'LDG', 'STG', 'STS', 'LDS', 'LDL', 'STL')) else op.split('.')[0]
        inst = int(float(r[ix['Instructions Executed']] or 0))
        samp = int(float(r[ix['# Samples']] or 0))
        by_op[base][0] += inst
        by_op[base][1] += samp
        for c in stall_cols:
            v = int(float(r[ix[c]] or 0))
            if v:
                by_op[base][2][c] += v
        tot_inst += inst
        tot_samp += samp
        lines.append((samp, inst, r[ix['Address']], src, {c: int(float(r[ix[c]] or 0)) for c in stall_cols}))
    print(f'total warp instructions {tot_inst}, samples {tot_samp}')
    print('--- by opcode (sorted by samples)')
    for op, (inst, samp, st) in sorted(by_op.items(), key=lambda kv: -kv[1][1])[:top]:
        s = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(f'{op:18s} inst {inst:10d} ({100.0 * inst / max(tot_inst, 1):5.1f}%)  samples {samp:7d} ({100.0 * samp / max(tot_samp, 1):5.1f}%)  '
              + ' '.join(f'{k[6:]}={v}' for k, v in s))
    print('--- hottest instructions')
    for samp, inst, addr, src, st in sorted(lines, key=lambda x: -x[0])[:top]:
        s = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print(f'{samp:6d} {inst:9d}  {src[:70]:70s} ' + ' '.join(f'{k[6:]}={v}' for k, v in s if v))
    agg = defaultdict(int)
    for _, _, _, _, st in lines:
        for k, v in st.items():
            agg[k] += v
    print('--- stall reasons overall')
    print(' '.join(f'{k[6:]}={v}' for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))


if __name__ == '__main__':
    main()
