#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table.

    python scripts/summarize_launches.py gpurun_out/launches.csv [--anchor text_maxagg_fwd] [--top 40]

One steady-state training step is the span between the last two launches of the anchor kernel
(one launch per step).  Per-launch times are cold-cache and serialised: read SHARES, not absolutes.
"""
import argparse
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r'<.*', '', name)          # drop template arguments
    name = re.sub(r'\(.*', '', name)          # drop parameter lists
    name = name.replace('at::', '').replace('void ', '')
    return name.strip()[:70]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('csv')
    ap.add_argument('--anchor', default='text_maxagg_fwd')
    ap.add_argument('--top', type=int, default=40)
    a = ap.parse_args()
    rows = []
    with open(a.csv, newline='') as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(unit, 1)
        rows.append((r['Kernel Name'], ns, r['Grid Size'], r['Block Size']))
    idx = [i for i, r in enumerate(rows) if a.anchor in r[0]]
    if len(idx) >= 2:
        lo, hi = idx[-2], idx[-1]
        span = rows[lo:hi]
        note = "one step = launches [%d, %d) between the last two `%s` launches" % (lo, hi, a.anchor)
    else:
        span = rows
        note = "anchor not found twice: whole list"
    agg = OrderedDict()
    for name, ns, grid, block in span:
        k = short(name)
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + ns)
    total = sum(t for _, t in agg.values())
    print("%s: %d launches, %.2f ms of kernel time\n" % (note, len(span), total / 1e6))
    print("| kernel | launches | us | share |\n|---|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
        print("| `%s` | %d | %.1f | %.1f%% |" % (k, n, t / 1e3, 100 * t / total))
    ours = sum(t for k, (n, t) in agg.items() if 'mgnns::' in k or k.startswith('tc::'))
    print("\nhand-written kernels (`mgnns::*`, incl. `mgnns::tc::*`): %.1f%% of the step's kernel time" % (100 * ours / total))


if __name__ == '__main__':
    sys.exit(main())
