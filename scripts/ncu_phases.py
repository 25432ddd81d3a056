#!/usr/bin/env python
"""Per-phase stall profile of one kernel from an .ncu-rep captured with --import-source on: splits the SASS at barriers
(BAR / WARPSYNC) and prints each segment's share of the warp-stall samples with its top stall reasons, then the hottest
instructions.  Usage: python scripts/ncu_phases.py prof.ncu-rep [min_samples_for_hot_list]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
hot = int(sys.argv[2]) if len(sys.argv) > 2 else 300
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1])
h = rows[1]
ia, isrc = h.index('Warp Stall Sampling (All Samples)'), h.index('Source')
names = [n for n in h if n.startswith('stall_') and '(' not in n]
idx = {n: h.index(n) for n in names}
body = [r for r in rows[2:] if len(r) > ia]
tot = sum(int(r[ia]) for r in body)
segs, cur = [], [0, 0, {}, '']
for r in body:
    toks = r[isrc].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    cur[0] += int(r[ia]); cur[1] += 1
    for n in names:
        cur[2][n] = cur[2].get(n, 0) + int(r[idx[n]])
    if op.startswith('BAR') or op.startswith('WARPSYNC'):
        cur[3] = op
        segs.append(cur); cur = [0, 0, {}, '']
segs.append(cur)
print('total samples', tot)
for s in segs:
    top = sorted(s[2].items(), key=lambda kv: -kv[1])[:6]
    print('%5.1f%% n_instr=%4d ends=%-10s' % (100 * s[0] / tot, s[1], s[3]), ' '.join('%s=%.1f' % (k.replace('stall_', ''), 100 * v / tot) for k, v in top if v / tot > 0.004))
print('--- hottest instructions')
for n, r in enumerate(body):
    if int(r[ia]) >= hot:
        print(n, r[ia], r[isrc].strip()[:80], ' '.join('%s=%s' % (k.replace('stall_', ''), r[i]) for k, i in idx.items() if int(r[i]) > hot // 6))
