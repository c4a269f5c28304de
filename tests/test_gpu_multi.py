"""Multi-GPU tests (need >= 2 visible GPUs; skipped otherwise — `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).

The one collective of the training step is the gradient all-reduce.  These tests run it over NCCL with
model.branch_streams = True, the configuration in which the AccumulateGrad hooks of one bucket fire on different CUDA
streams (mgnns_b200/ddp.py: every copy records an event, the collective waits for all of them)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, 'tests'))
import torch, torch.distributed as dist
import mgnns_test_helpers as H
from mgnns_b200 import synth
from mgnns_b200.ddp import GradientAllReducer
from mgnns_b200.api.multi_gcn import IdentityTrunk, Multi_GCN_Multihead_Att
from mgnns_b200.api.text_gcn import Model as TextModel
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank); dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
cfg = dict(H.MODEL_CFG, B=16, V=300, seed=21)
emap, count = synth.synthetic_edge_map(cfg['V'], seed=21, docs=400)
def build():
    vocab = ['PAD', 'UNK'] + ['w%%d' %% i for i in range(2, cfg['V'])]
    tm = TextModel(7, 300, vocab, cfg['ngram'], 0.5, count, emap, pmi=torch.zeros(count, 1))
    opt = dict(emb_path='', bidirectional=True, hidden_size=150, emb_size=300, num_layers=2, dropout=0.5, emb_type='random',
               vocab_size=cfg['V'], stack_num=2, n_head=4, d_kv=128, is_regu=False)
    m = Multi_GCN_Multihead_Att(opt, 7, tm, IdentityTrunk(), IdentityTrunk(), 80, 365, object_t=0.4, place_t=0.3, in_channel=300,
                                object_adj_file=synth.adj_dict('object'), place_adj_file=synth.adj_dict('place'))
    synth.fill_parameters(m, seed=21)
    return m.to(dev).eval()            # eval: no dropout, so the sharded and the whole-batch runs are comparable
text, lens, mask, fo, fp, oinp, pinp, labels = H.model_inputs(cfg)
def run(model, sl):
    logits = model(text[sl].to(dev), lens[sl], mask[sl].to(dev), fo[sl].to(dev), fp[sl].to(dev), oinp[sl].to(dev), pinp[sl].to(dev))
    return torch.nn.functional.cross_entropy(logits, labels[sl].to(dev))
per = cfg['B'] // world
mine = slice(rank * per, (rank + 1) * per)
model = build(); model.branch_streams = True
red = GradientAllReducer(model, bucket_bytes=4 << 20)
for step in range(3):                  # step 0 builds the buckets, steps 1-2 take the hook path
    model.zero_grad(set_to_none=True)
    run(model, mine).backward()
    red.finish()
torch.cuda.synchronize()
ref = build(); ref.branch_streams = False
run(ref, slice(0, cfg['B'])).backward()       # mean over the whole batch == mean of the per-rank means (equal shards)
worst = 0.0
for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
    if q.grad is None:
        assert p.grad is None, n
        continue
    d = (p.grad - q.grad).abs().max().item()
    s = q.grad.abs().max().item()
    assert d <= 2e-5 + 2e-4 * s, (n, d, s)
    worst = max(worst, d)
assert red.payload_bytes() > 1e7
dist.barrier()
if rank == 0:
    print('ok', worst)
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_bucketed_allreduce_with_branch_streams_matches_whole_batch_gradients(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {'root': ROOT})
    env = dict(os.environ, MASTER_ADDR='127.0.0.1')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                          '--master-addr', '127.0.0.1', '--master-port', '29613', str(script)],
                         capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0 and 'ok' in out.stdout, (out.stdout[-1000:], out.stderr[-3000:])


P2P_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, 'tests'))
import torch, torch.distributed as dist
from mgnns_b200.p2p import PeerAllReduce
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank); dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
for numel in (4, 1000, 3_000_004, 24_858_240):            # tiny, ragged slices, odd float4 count, the cfg-4 payload
    peer = PeerAllReduce(numel, dev, ctas=32)
    for it in range(4):                                     # repeated launches: the device-side epochs advance
        g = torch.Generator(device='cpu').manual_seed(1000 * it + rank)
        x = torch.randn(numel, generator=g).to(dev)
        peer.flat.copy_(x)
        ref = x.double()
        parts = [torch.empty_like(ref) for _ in range(world)]
        dist.all_gather(parts, ref)
        want = (sum(parts) / world).float()                 # fp64 sum of the same inputs
        peer.all_reduce_(1.0 / world)
        torch.cuda.synchronize()
        peer.check()
        got = peer.flat.clone()
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-6), (numel, it, (got - want).abs().max().item())
        # bit-identical on every rank: each element is summed once, by its owner, in rank order
        allg = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(allg, got)
        assert all(torch.equal(allg[0], t) for t in allg), (numel, it)
    peer.close()
# inside a CUDA graph, replayed: the kernel carries its own epochs, no host state per replay
peer = PeerAllReduce(1_000_000, dev)
base = torch.full((1_000_000,), float(rank + 1), device=dev)
side = torch.cuda.Stream(device=dev); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    peer.flat.copy_(base); peer.all_reduce_(1.0)
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    peer.flat.copy_(base)
    peer.all_reduce_(1.0)
for _ in range(5):
    graph.replay()
torch.cuda.synchronize(); peer.check()
assert torch.equal(peer.flat, torch.full_like(base, float(world * (world + 1) // 2)))
dist.barrier()
if rank == 0:
    print('ok')
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_memory_allreduce_matches_fp64_sum_and_is_bit_identical_across_ranks(tmp_path):
    """mgnns_allreduce_p2p_f32 (one kernel over NVLink peer memory) on every visible GPU: mean of random buffers vs an
    fp64 sum, identical bits on all ranks, repeated launches, and replay inside a CUDA graph."""
    script = tmp_path / "p2p_worker.py"
    script.write_text(P2P_WORKER % {'root': ROOT})
    n = min(torch.cuda.device_count(), 8)
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=%d' % n,
                          '--master-addr', '127.0.0.1', '--master-port', '29617', str(script)],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, MASTER_ADDR='127.0.0.1'))
    assert out.returncode == 0 and 'ok' in out.stdout, (out.stdout[-1000:], out.stderr[-3000:])


GRAPH_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, 'tests'))
os.environ['MGNNS_P2P_ALLREDUCE'] = sys.argv[1]
import torch, torch.distributed as dist
import mgnns_test_helpers as H
from mgnns_b200 import synth
from mgnns_b200.graph_step import GraphedTrainStep
from mgnns_b200.api.multi_gcn import IdentityTrunk, Multi_GCN_Multihead_Att
from mgnns_b200.api.text_gcn import Model as TextModel
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank); dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
cfg = dict(H.MODEL_CFG, B=16, V=300, seed=33)
emap, count = synth.synthetic_edge_map(cfg['V'], seed=33, docs=400)
vocab = ['PAD', 'UNK'] + ['w%%d' %% i for i in range(2, cfg['V'])]
tm = TextModel(7, 300, vocab, cfg['ngram'], 0.5, count, emap, pmi=torch.zeros(count, 1))
opt_cfg = dict(emb_path='', bidirectional=True, hidden_size=150, emb_size=300, num_layers=2, dropout=0.5, emb_type='random',
               vocab_size=cfg['V'], stack_num=2, n_head=4, d_kv=128, is_regu=False)
m = Multi_GCN_Multihead_Att(opt_cfg, 7, tm, IdentityTrunk(), IdentityTrunk(), 80, 365, object_t=0.4, place_t=0.3, in_channel=300,
                            object_adj_file=synth.adj_dict('object'), place_adj_file=synth.adj_dict('place'))
synth.fill_parameters(m, seed=33)
m = m.to(dev).eval(); m.branch_streams = True
text, lens, mask, fo, fp, oinp, pinp, labels = H.model_inputs(cfg)
per = cfg['B'] // world
sl = slice(rank * per, (rank + 1) * per)
batch = dict(text=text[sl].to(dev), lens=lens[sl].clone(), mask=mask[sl].to(dev), fo=fo[sl].to(dev), fp=fp[sl].to(dev),
             oinp=oinp[sl].to(dev), pinp=pinp[sl].to(dev), labels=labels[sl].to(dev))
o = torch.optim.Adam(m.get_config_optim(1e-3, 0.1), lr=1e-3, weight_decay=1e-5, eps=1e-2, capturable=True, fused=True)
g = GraphedTrainStep(m, o, torch.nn.CrossEntropyLoss(), batch, clip_norm=10.0, world_size=world, warmup=1, plan_capacity=per * 100,
                     flat_optimizer=True)
assert g.use_p2p == (sys.argv[1] == '1')
losses = [float(g.replay()) for _ in range(3)]
torch.cuda.synchronize()
if g.use_p2p:
    g._fg.peer.check()
# every rank holds the same parameters after the same all-reduced updates
chk = torch.stack([p.detach().double().sum() for p in m.parameters()])
allc = [torch.empty_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
assert all(torch.equal(allc[0], t) for t in allc), 'ranks diverged'
if rank == 0:
    torch.save({'chk': chk.cpu(), 'losses': losses}, sys.argv[2])
    print('ok')
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_graphed_step_with_peer_memory_allreduce_matches_nccl_path(tmp_path):
    """Four data-parallel training steps (warm-up + 3 replays) with the all-reduce as a kernel node inside ONE captured
    graph vs the NCCL all-reduce between two captured halves: same parameters (to summation order), ranks in sync."""
    script = tmp_path / "graph_worker.py"
    script.write_text(GRAPH_WORKER % {'root': ROOT})
    res = {}
    for mode, port in (('1', '29621'), ('0', '29623')):
        out_file = tmp_path / ("res%s.pt" % mode)
        out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                              '--master-addr', '127.0.0.1', '--master-port', port, str(script), mode, str(out_file)],
                             capture_output=True, text=True, timeout=900, env=dict(os.environ, MASTER_ADDR='127.0.0.1'))
        assert out.returncode == 0 and 'ok' in out.stdout, (mode, out.stdout[-1000:], out.stderr[-3000:])
        res[mode] = torch.load(out_file)
    a, b = res['1'], res['0']
    assert torch.allclose(torch.tensor(a['losses']), torch.tensor(b['losses']), rtol=2e-4)
    assert torch.allclose(a['chk'], b['chk'], rtol=1e-4, atol=1e-4)
