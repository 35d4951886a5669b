#!/usr/bin/env python
"""Join an ncu source-page CSV (SASS order) with nvdisasm -g line info to attribute executed
instructions and stall samples to CUDA source lines.
usage: line_profile.py <ncu_source.csv> <nvdisasm -g -c output> <function prefix> <units>"""
import collections
import csv
import re
import sys

src_csv, sass_txt, fn_prefix, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
txt = open(sass_txt).read().split('\n')
start = [i for i, l in enumerate(txt) if l.startswith('.text.' + fn_prefix)][0]
end = len(txt)
for i in range(start + 1, len(txt)):
    if txt[i].startswith('//--------------------- .text.'):
        end = i
        break
cur, instr = None, []
for l in txt[start:end]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        instr.append((cur, m.group(2)))
rows = list(csv.reader(open(src_csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]
ie, isamp = hdr.index('Instructions Executed'), hdr.index('# Samples')
data = [r for r in rows[hi + 1:] if len(r) > ie]
assert len(data) == len(instr), (len(data), len(instr))
agg, aggs, opc = collections.Counter(), collections.Counter(), collections.Counter()
for (loc, op), r in zip(instr, data):
    try:
        c, s = int(r[ie]), int(r[isamp])
    except ValueError:
        continue
    agg[loc] += c
    aggs[loc] += s
    o = op.split()
    o = o[1] if o[0].startswith('@') else o[0]
    opc[o.split('.')[0]] += c
tot, ts = sum(agg.values()), sum(aggs.values())
print('warp instructions per unit', tot / units)
cache = {}
for k, v in agg.most_common(int(sys.argv[5]) if len(sys.argv) > 5 else 30):
    line = ''
    if k:
        try:
            if k[0] not in cache:
                cache[k[0]] = open('roboticmanipulators_b200/csrc/' + k[0]).read().split('\n')
            line = cache[k[0]][k[1] - 1].strip()[:66]
        except OSError:
            pass
    print(f"{v / units:7.1f} {100 * v / tot:5.1f}% smp {100 * aggs[k] / ts:5.1f}%  {k}  {line}")
print({k: round(v / units, 1) for k, v in opc.most_common(18)})
