"""CPU/GPU breakdown of one bench training step (run on the GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import bench
from mgnns_b200 import synth
dev = torch.device('cuda', 0)
emap, count = synth.synthetic_edge_map(bench.VOCAB, seed=0, docs=20000)
model = bench.build_model(dev, emap, count).train()
opt = torch.optim.Adam(model.get_config_optim(5e-5, 0.1), lr=5e-5, weight_decay=1e-5)
B = int(os.environ.get('B', 512))
d = bench.to_device(bench.host_batch(B, 0), dev, B)
crit = torch.nn.CrossEntropyLoss()
def step():
    opt.zero_grad(set_to_none=True)
    logits = model(d['text'], d['lens'], d['mask'], d['fo'], d['fp'], d['oinp'], d['pinp'])
    loss = crit(logits, d['labels']); loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
# phase timing with syncs
def timed(fn):
    torch.cuda.synchronize(); t=time.perf_counter(); r=fn(); cpu=time.perf_counter()-t; torch.cuda.synchronize(); return r, cpu, time.perf_counter()-t
for _ in range(2):
    opt.zero_grad(set_to_none=True)
    logits, c1, w1 = timed(lambda: model(d['text'], d['lens'], d['mask'], d['fo'], d['fp'], d['oinp'], d['pinp']))
    loss = crit(logits, d['labels'])
    _, c2, w2 = timed(lambda: loss.backward())
    _, c3, w3 = timed(lambda: torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0))
    _, c4, w4 = timed(lambda: opt.step())
    print('fwd cpu %.1f wall %.1f | bwd cpu %.1f wall %.1f | clip cpu %.1f wall %.1f | adam cpu %.1f wall %.1f (ms)' % tuple(1e3*x for x in (c1,w1,c2,w2,c3,w3,c4,w4)))
# sub-phase timing inside forward
def seg(name, fn):
    r, c, w = timed(fn); print('  %-28s cpu %.2f wall %.2f ms' % (name, c*1e3, w*1e3)); return r
with torch.no_grad():
    tf = seg('text_features', lambda: model.text_features(d['text']))
    tb = seg('lstm bank', lambda: model.get_text_memory_bank(d['text'], d['lens'], True)[0])
    ob = seg('imgbank obj', lambda: model._img_bank(d['fo'], model.liner_img_object))
    q = model._query()
    oa = seg('label channel obj', lambda: model._label_channel(ob[1], d['oinp'], 'object_A', model.object_attention, model.object_linear_5, model.object_x_linear, q))
    x = oa
    for i, layer in enumerate(model.img_object_text_multi_head_att):
        x = seg('mha layer text %d' % i, lambda: layer(q=x, k=tb, v=tb, mask=d['mask'])[0])
    x = tf
    for i, layer in enumerate(model.text_img_object_multi_head_att):
        x = seg('mha layer img %d' % i, lambda: layer(q=x, k=ob[0], v=ob[0])[0])
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
