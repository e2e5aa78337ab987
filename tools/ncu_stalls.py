#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel in an .ncu-rep (source page):  tools/ncu_stalls.py rep [kernel-index] [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; kid = sys.argv[2] if len(sys.argv) > 2 else "1"; N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][:2])
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[idx['# Samples']].isdigit()]
tot = sum(int(r[idx['# Samples']]) for r in data)
print('total samples', tot, 'instructions', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[idx[s]]) for r in data) for s in stalls}
print('stall totals:', sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:N]:
    st = sorted([(int(r[idx[s]]), s[6:]) for s in stalls], reverse=True)[:2]
    print(r[idx['# Samples']].rjust(6), r[idx['Instructions Executed']].rjust(8), r[idx['Source']][:80].ljust(80), st)
