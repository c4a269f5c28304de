"""Extract per-launch DRAM traffic of our kernels from `ncu --set full` reports into profiles/ncu_traffic.json
(bench.py's roofline.traffic) and print a markdown table.

    python scripts/extract_traffic.py gpurun_out/a.ncu-rep[:alias=kernel-regex,...] ...
Each report may hold several kernels; `alias=regex` pairs name the bench.py timer a kernel belongs to."""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'lts__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__m_xbar2l1tex_read_bytes.sum']
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}


def rows_of(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rd = csv.reader(out.splitlines())
    hdr, units = next(rd), next(rd)
    for row in rd:
        d = {'name': row[hdr.index('Kernel Name')], 'grid': row[hdr.index('Grid Size')], 'block': row[hdr.index('Block Size')]}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                d[w] = float(row[i].replace(',', '')) * UNIT.get(units[i], 1.0)
        yield d


def main():
    table, traffic = [], {}
    for arg in sys.argv[1:]:
        rep, _, spec = arg.partition(':')
        aliases = [a.split('=', 1) for a in spec.split(',') if a]
        seen = set()
        for d in rows_of(rep):
            key = (d['name'], d['grid'])
            if key in seen:
                continue
            seen.add(key)
            alias = next((a for a, rx in aliases if re.search(rx, d['name'])), None)
            bytes_ = d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
            table.append((alias or '-', d['name'].split('(')[0][-48:], d['grid'], d['block'], d))
            if alias and alias not in traffic:
                traffic[alias] = {'kernel': d['name'].split('(')[0], 'dram_bytes_per_launch': bytes_,
                                  'dram_read': d.get('dram__bytes_read.sum'), 'dram_write': d.get('dram__bytes_write.sum'),
                                  'us_under_ncu': d.get('gpu__time_duration.sum'), 'report': os.path.basename(rep)}
    print("| timer | kernel | grid x block | time (us) | DRAM read (MB) | DRAM write (MB) | DRAM % | tensor pipe % | "
          "issue active % | warps active % | regs | L2 hit % | L1 data pipe % | L1 hit % | L2->L1 (GB) |\n|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for alias, name, grid, block, d in table:
        g = lambda k, s=1.0: ('%.1f' % (d[k] / s)) if k in d else '-'
        print("| %s | `%s` | %s x %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |" % (
            alias, name, grid, block, g(WANT[0]), g(WANT[1], 1e6), g(WANT[2], 1e6), g(WANT[3]), g(WANT[4]), g(WANT[8]),
            g(WANT[5]), g(WANT[6]), g(WANT[7]), g(WANT[9]), g(WANT[10]), g(WANT[11], 1e9)))
    if traffic:
        path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
        old = json.load(open(path)) if os.path.exists(path) else {}
        old.update(traffic)
        json.dump(old, open(path, 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
