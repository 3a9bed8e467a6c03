#!/usr/bin/env python
"""ptxas does not fit the code after `setmaxnreg.dec N` into N registers by itself: scans the SASS of an object file and
reports, per kernel, the highest register index used between USETMAXREG.DEALLOC and the first USETMAXREG.TRY_ALLOC (the
branch of the warps that released registers).  Exit code 1 when a kernel exceeds its own limit.
    python tools/check_setmaxnreg.py cross-scale-mae_b200/csmae_b200/lib/obj/attention_tc.o"""
import re
import subprocess
import sys


def main():
    bad = 0
    for obj in sys.argv[1:]:
        sass = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
        name, limit, inside, top = None, None, False, -1
        for line in sass.splitlines():
            m = re.search(r'Function : (\S+)', line)
            if m:
                name, limit, inside, top = m.group(1), None, False, -1
                continue
            if 'USETMAXREG.DEALLOC' in line and not inside and limit is None:
                limit = int(re.search(r'0x([0-9a-f]+) ;', line).group(1), 16)
                inside = True
                continue
            if 'USETMAXREG.TRY_ALLOC' in line and inside:
                inside = False
                short = re.sub(r'^_ZN\d+_GLOBAL__N__[0-9a-f_]+cu_[0-9a-f]+', '', name)[:60]
                ok = top < limit
                bad += 0 if ok else 1
                print(f'{"ok " if ok else "BAD"} {short:60s} dealloc to {limit:3d}, highest register used R{top}')
            if inside:
                for r in re.findall(r'\bR(\d+)\b', line.split('/*')[1] if line.count('/*') > 1 else line):
                    top = max(top, int(r))
    return 1 if bad else 0


if __name__ == '__main__':
    sys.exit(main())
