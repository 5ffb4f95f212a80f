#!/usr/bin/env python3
"""Per-kernel counts of the SASS mnemonics that tell a Blackwell-native kernel from a recompiled one
(B200_PROFILING.md "What proves a Blackwell-native kernel"): tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM,
TMA -> UTMALDG/UTMASTG/UTMAREDG/UBLKCP, mma.sync -> HMMA, cp.async -> LDGSTS.
    python scripts/sass_opcodes.py [lib.so] > profiles/sass_opcodes.md"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else 'argsim_b200/libargsim_b200.so'
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
pats = [('UTC*MMA', r'\bUTC[A-Z]*MMA'), ('LDTM', r'\bLDTM'), ('STTM', r'\bSTTM'), ('UTMALDG', r'\bUTMALDG'),
        ('UTMASTG', r'\bUTMASTG'), ('UTMAREDG', r'\bUTMAREDG'), ('UBLKCP', r'\bUBLKCP'), ('HMMA', r'\bHMMA'),
        ('LDGSTS', r'\bLDGSTS'), ('SYNCS', r'\bSYNCS'), ('MUFU', r'\bMUFU'), ('total', r'^\s+/\*[0-9a-f]{4}\*/')]
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for name, p in pats:
        if re.search(p, line):
            counts[cur][name] += 1


def demangle(n):
    try:
        return subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
    except Exception:
        return n


print('# SASS opcode census of `%s` (cuobjdump -sass, sm_100a)\n' % lib)
print('tcgen05.mma = `UTC*MMA`, tcgen05.ld/st = `LDTM`/`STTM`, TMA = `UTMALDG`/`UTMASTG`/`UTMAREDG`, mma.sync = `HMMA`, '
      'cp.async = `LDGSTS`, mbarrier = `SYNCS`.\n')
cols = [n for n, _ in pats]
print('| kernel | ' + ' | '.join(cols) + ' |')
print('|---|' + '---:|' * len(cols))
for k, c in counts.items():
    name = demangle(k)
    name = re.sub(r'\(anonymous namespace\)::', '', name)
    name = re.sub(r'\(.*', '', name)
    name = re.sub(r'^void ', '', name)
    print('| `%s` | ' % name[:90] + ' | '.join(str(c.get(n, 0)) for n in cols) + ' |')
