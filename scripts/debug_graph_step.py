"""Bisect graph-vs-eager differences of one training step (eval mode): prints the parameters whose gradient / value
differ most after ONE step, for the graph step with and without branch streams / deferred weight gradients."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import mgnns_test_helpers as H
from mgnns_b200 import ops, synth
from mgnns_b200.graph_step import GraphedTrainStep
from test_gpu_parity import build_model

dev = torch.device('cuda', 0)
cfg = dict(H.MODEL_CFG, B=16, V=300, seed=51)
emap, count = synth.synthetic_edge_map(cfg['V'], seed=51, docs=500)
text, lens, mask, fo, fp, oinp, pinp, labels = H.model_inputs(cfg)
b = dict(text=text.to(dev), lens=lens, mask=mask.to(dev), fo=fo.to(dev), fp=fp.to(dev), oinp=oinp.to(dev), pinp=pinp.to(dev),
         labels=labels.to(dev))
crit = torch.nn.CrossEntropyLoss()

def fresh():
    m = build_model(dev, cfg, emap, count).eval()
    o = torch.optim.Adam(m.get_config_optim(1e-3, 0.1), lr=1e-3, weight_decay=1e-5, eps=1e-2, capturable=True, fused=True)
    return m, o

m_e, o_e = fresh()
o_e.zero_grad(set_to_none=True)
loss = crit(m_e(b['text'], b['lens'], b['mask'], b['fo'], b['fp'], b['oinp'], b['pinp']), b['labels'])
loss.backward()
CLIP = float(os.environ.get('CLIP', '10.0'))
pre = torch.sqrt(sum((p.grad.detach().double() ** 2).sum() for p in m_e.parameters() if p.grad is not None)).item()
tn = torch.nn.utils.clip_grad_norm_(m_e.parameters(), CLIP)
print("eager pre-clip total norm", pre, "returned", float(tn))
ge = {n: p.grad.detach().clone() for n, p in m_e.named_parameters() if p.grad is not None}
o_e.step()
pe = {n: p.detach().clone() for n, p in m_e.named_parameters()}
print("eager loss", loss.item(), "grads", len(ge))

for streams, defer in ((False, False), (True, True)):
    m_g, o_g = fresh()
    m_g.branch_streams = streams
    if not defer:
        orig = ops.defer_weight_grads
        ops.defer_weight_grads = lambda enabled: orig(False)
    static = {k: (v.clone() if torch.is_tensor(v) and k != 'lens' else v) for k, v in b.items()}
    p0 = {n: p.detach().clone() for n, p in m_g.named_parameters()}
    g = GraphedTrainStep(m_g, o_g, crit, static, clip_norm=CLIP, world_size=1, warmup=1, plan_capacity=1600)
    if not defer:
        ops.defer_weight_grads = orig
    with torch.no_grad():
        for n, p in m_g.named_parameters():
            p.copy_(p0[n])
    for st in o_g.state.values():
        for k, v in st.items():
            if torch.is_tensor(v):
                v.zero_()
    l = g.replay().item()
    torch.cuda.synchronize()
    gg = {n: p.grad.detach() for n, p in m_g.named_parameters() if p.grad is not None}
    print("\nstreams=%s defer=%s: graph loss %.6f, grads %d (eager %d), missing %s extra %s" % (
        streams, defer, l, len(gg), len(ge), sorted(set(ge) - set(gg))[:5], sorted(set(gg) - set(ge))[:5]))
    dg = sorted(((float((gg[n] - ge[n]).abs().max() / (ge[n].abs().max() + 1e-12)), n) for n in gg if n in ge), reverse=True)[:6]
    print("  worst relative grad diffs:", [(round(d, 6), n) for d, n in dg])
    dp = sorted(((float((p.detach() - pe[n]).abs().max()), n) for n, p in m_g.named_parameters()), reverse=True)[:6]
    print("  worst abs param diffs after step:", [("%.2e" % d, n) for d, n in dp])
    steps = [float(st['step']) for st in o_g.state.values() if 'step' in st][:3]
    print("  optimizer step counters:", steps)
    post = torch.sqrt(sum((x.double() ** 2).sum() for x in gg.values())).item()
    print("  graph post-clip total norm", post, " eager post-clip", torch.sqrt(sum((x.double() ** 2).sum() for x in ge.values())).item())
