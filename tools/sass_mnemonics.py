#!/usr/bin/env python
"""SASS mnemonic counts per kernel of the in-tree library (no GPU needed: `cuobjdump -sass` + `c++filt`).
usage: python tools/sass_mnemonics.py > profiles/<round>_sass_mnemonics.md
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk,
HMMA = mma.sync, LDGSTS = cp.async, UCGABAR_ARV = barrier.cluster.arrive, SYNCS = mbarrier ops, ATOMG / REDG = global atomics with / without a return value."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'valle2_b200', 'lib', 'libvalle_b200.so')
COLS = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'HMMA', 'LDGSTS', 'UCGABAR_ARV', 'SYNCS', 'ATOMG', 'REDG', 'MUFU']


def main() -> int:
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    kernels, name = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m and name:
            op = m.group(1)
            kernels[name]['_n'] += 1
            for c in COLS:
                if op == c or op.startswith(c + '.') or (c == 'SYNCS' and op.startswith('SYNCS')) or (c == 'MUFU' and op.startswith('MUFU')):
                    kernels[name][c] += 1
    names = list(kernels)
    dem = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True, check=True).stdout.splitlines()
    rows = []
    for raw, d in zip(names, dem):
        d = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', d)
        d = re.sub(r'\(.*$', '', d)                       # drop the parameter list
        rows.append((d, kernels[raw]))
    rows.sort(key=lambda r: r[0])
    print('# SASS mnemonics per kernel of valle2_b200/lib/libvalle_b200.so\n')
    print('`python tools/sass_mnemonics.py` (`cuobjdump -sass` of the in-tree library; nvcc 12.9, `-gencode arch=compute_100a,code=sm_100a`).')
    print('UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk,')
    print('HMMA = mma.sync, LDGSTS = cp.async, UCGABAR_ARV = barrier.cluster.arrive, SYNCS = mbarrier ops, ATOMG / REDG = global atomics with / without a return value,')
    print('MUFU = special-function unit.  Kernels with none of the first nine are plain SIMT row kernels.\n')
    print('| kernel | instructions | ' + ' | '.join(COLS) + ' |')
    print('|---|---:|' + '---:|' * len(COLS))
    for d, c in rows:
        print(f'| `{d}` | {c["_n"]} | ' + ' | '.join(str(c[k]) if c[k] else '' for k in COLS) + ' |')
    return 0


if __name__ == '__main__':
    sys.exit(main())
