"""Kernel timeline of the CUDA-graph training step (torch profiler / CUPTI activity records, no ncu replay):
which kernels are on the GPU when, how much of the step has 1, 2, 3... kernels in flight, and which kernels
run ALONE (the serial part of the step).  Run on the GPU box: python scripts/timeline_step.py [B]"""
import json
import os
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch

import bench
from mgnns_b200 import synth
from mgnns_b200.graph_step import GraphedTrainStep

dev = torch.device('cuda', 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
emap, count = synth.synthetic_edge_map(bench.VOCAB, seed=0, docs=20000)
model = bench.build_model(dev, emap, count).train()
model.branch_streams = os.environ.get('MGNNS_BRANCH_STREAMS', '1') == '1'
opt = torch.optim.Adam(model.get_config_optim(5e-5, 0.1), lr=5e-5, weight_decay=1e-5, capturable=True, fused=True)
d = bench.to_device(bench.host_batch(B, 0), dev, B)
g = GraphedTrainStep(model, opt, torch.nn.CrossEntropyLoss(), d, clip_norm=10.0, world_size=1, warmup=2,
                     flat_optimizer=None if os.environ.get('MGNNS_TORCH_OPT') == '1' else True)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.replay()                       # exactly one step in the trace
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), 'trace.json')
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset') and 'dur' in e]
ev.sort(key=lambda e: e['ts'])
step = ev
t0 = step[0]['ts']
span = max(e['ts'] + e['dur'] for e in step) - t0
print("one replay: %d GPU activities over %.3f ms on %d streams" % (len(step), span / 1e3, len({e['args'].get('stream') for e in step})))
# sweep line: concurrency histogram and per-kernel "alone" time
pts = []
for i, e in enumerate(step):
    pts.append((e['ts'], 1, i))
    pts.append((e['ts'] + e['dur'], -1, i))
pts.sort()
active = set()
conc = defaultdict(float)
alone = defaultdict(float)
total = defaultdict(float)
prev = pts[0][0]
for t, kind, i in pts:
    dt = t - prev
    if dt > 0:
        conc[len(active)] += dt
        if len(active) == 1:
            alone[short := step[next(iter(active))]['name'][:60]] += dt
    prev = t
    if kind == 1:
        active.add(i)
    else:
        active.discard(i)
for e in step:
    total[e['name'][:60]] += e['dur']
print("time with N activities in flight: " + ", ".join("%d: %.2f ms" % (k, v / 1e3) for k, v in sorted(conc.items())))
print("\n%-62s %10s %10s" % ("kernel", "total us", "alone us"))
for k, v in sorted(total.items(), key=lambda kv: -alone.get(kv[0], 0))[:28]:
    print("%-62s %10.1f %10.1f" % (k, v, alone.get(k, 0.0)))

# compact Gantt of the long activities (>= 40 us) and of every stream's busy span
print("\nstart_ms  end_ms  stream  kernel")
min_dur = 0 if os.environ.get('MGNNS_TIMELINE_ALL') == '1' else 40
for e in step:
    if e['dur'] >= min_dur:
        print("%7.3f %7.3f  %6s  %s" % ((e['ts'] - t0) / 1e3, (e['ts'] + e['dur'] - t0) / 1e3, e['args'].get('stream'), e['name'][:70]))
by_stream = defaultdict(list)
for e in step:
    by_stream[e['args'].get('stream')].append(e)
print("\nstream: first start, last end, busy ms, activities")
for sid, evs in sorted(by_stream.items(), key=lambda kv: kv[1][0]['ts']):
    print("%6s  %7.3f %7.3f  %6.3f  %d" % (sid, (evs[0]['ts'] - t0) / 1e3, (max(x['ts'] + x['dur'] for x in evs) - t0) / 1e3,
                                             sum(x['dur'] for x in evs) / 1e3, len(evs)))
