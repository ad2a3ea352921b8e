"""Per-kernel SASS mnemonic counts of the built library (cuobjdump -sass): which kernels are tcgen05 / TMEM / bulk-copy
code.  python tools/sass_summary.py > profiles/<name>.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'kgdet_b200', '_lib', 'libkgdet_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], stdout=subprocess.PIPE, text=True).stdout
WATCH = ('UTCHMMA', 'UTCQMMA', 'LDTM', 'UBLKCP', 'UBLKRED', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'UTCBAR', 'HMMA', 'REDG',
         'ATOMG', 'ATOMS', 'MUFU', 'HFMA2', 'LDGSTS', 'MATCH', 'REDUX', 'SYNCS')
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = counts.setdefault(m.group(1), collections.Counter())
        continue
    if cur is None:
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m:
        op = m.group(1)
        for w in WATCH:
            if op.startswith(w):
                cur[w] += 1
                break
print('# SASS mnemonic counts per kernel of kgdet_b200/_lib/libkgdet_b200.so (cuobjdump -sass, sm_100a)')
print('# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, UBLKRED = cp.reduce.async.bulk, UTMALDG = '
      'cp.async.bulk.tensor (TMA), UTCBAR = tcgen05.commit, HMMA = legacy mma.sync (none expected)')
tot = collections.Counter()
for name, c in counts.items():
    tot.update(c)
    print('%-110s %s' % (name[:110], ' '.join('%s=%d' % kv for kv in sorted(c.items()))))
print('# total: ' + ' '.join('%s=%d' % kv for kv in sorted(tot.items())))
