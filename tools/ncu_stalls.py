#!/usr/bin/env python3
"""Per-instruction stall summary of an .ncu-rep captured with --import-source on (read here, no GPU needed).
usage: tools/ncu_stalls.py <report.ncu-rep>"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
print('total samples', tot)
keys = [h for h in hdr if h.startswith('stall_') and '(' not in h]
for key in keys:
    s = sum(int(r[ix[key]]) for r in data)
    if s: print(f'  {key:28s} {s:9d} {100*s/tot:5.1f} %')
for key in ('stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_barrier'):
    print('top', key)
    for r in sorted(data, key=lambda r: -int(r[ix[key]]))[:5]:
        print('   ', r[0][-6:], r[1][:70], r[ix[key]], 'exec', r[ix['Instructions Executed']])
print('sync / copy instructions')
for i, r in enumerate(data):
    if any(k in r[1] for k in ('TRYWAIT', 'ATOMS', 'BAR.SYNC', 'UBLKCP', 'UTMALDG')) and int(r[ix['Instructions Executed']]) > 0:
        print('   ', i, r[0][-6:], r[1][:60], 'samples', r[ix['# Samples']], 'exec', r[ix['Instructions Executed']])
