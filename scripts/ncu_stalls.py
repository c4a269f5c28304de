"""Per-kernel stall breakdown and the hottest SASS lines from an ncu source-page CSV
(ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-count 1 > src.csv; python scripts/ncu_stalls.py src.csv)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if r and r[0] == 'Kernel Name':
        break                                    # next launch in the same file
    if len(r) == len(hdr) and r[0] != 'Address':
        data.append(r)
tot = sum(int(r[ix['# Samples']]) for r in data)
print('total samples', tot, 'instructions', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    print('  %-24s %8d  %5.1f %%' % (s, v, 100.0 * v / max(tot, 1)))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:n]:
    st = {s: int(r[ix[s]]) for s in stalls if int(r[ix[s]]) > 0}
    main = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(r[ix['# Samples']].rjust(7), r[ix['Source']].strip()[:72].ljust(72), main)
