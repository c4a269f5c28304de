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
