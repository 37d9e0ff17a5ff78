"""Opcode mix and hottest instructions of an .ncu-rep captured with --import-source on: python tools/ncu_src.py <file> [kernel index]"""
import csv
import io
import subprocess
import sys
from collections import Counter

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) > 5]
ia, isrc, isamp, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
tot_ex = sum(int(r[iex] or 0) for r in data)
tot_s = sum(int(r[isamp] or 0) for r in data)
print('total warp instructions', tot_ex, 'samples', tot_s)
c, cs = Counter(), Counter()
for r in data:
    t = r[isrc].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    c[op] += int(r[iex] or 0)
    cs[op] += int(r[isamp] or 0)
for op, n in c.most_common(22):
    print('%-10s exec %10d (%4.1f%%)  samples %6d (%4.1f%%)' % (op, n, 100 * n / tot_ex, cs[op], 100 * cs[op] / max(tot_s, 1)))
print('--- hottest instructions')
for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(r[ia][-5:], '%5s %9s' % (r[isamp], r[iex]), r[isrc][:100])
