"""CPU tests: compat aliases; world-size-2 gloo run of the bucketed gradient all-reducer."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_compat_aliases_resolve_reference_import_names():
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import mgnns_b200.compat as c; c.install()\n"
        "from models.Multi_GCN_Multihead_att_new import multi_gcn_multihead_att_model, GraphConvolution, Attention\n"
        "from models.Multi_GCN_Multihead_att import Multi_GCN_Multihead_Att\n"
        "from models.multi_head_att.submodules import MultiHeadAttention, PositionwiseFeedForward, LayerNorm\n"
        "from models.moudles import CoAttention, MyMultiHeadAttention, MyAnotherMultiHeadAttention\n"
        "from models.Text_GCN import Model\n"
        "from utils.pmi import cal_PMI\n"
        "from utils.util import gen_A, gen_adj\n"
        "from utils.vocab_new import get_vocab_list\n"
        "import dgl, word2vec, torchnet, apex\n"
        "print('ok')\n" % ROOT)
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith('ok'), out.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir('/root/reference/engine'), reason="reference checkout not present")
def test_reference_engine_and_entry_imports_unchanged_with_compat():
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import mgnns_b200.compat as c; c.install(reference_root='/root/reference')\n"
        "import importlib\n"
        "eng = importlib.import_module('engine.Multi_GCN_Multihead_Att_engine')\n"
        "assert hasattr(eng, 'GCNMultiClassEngine')\n"
        "import models.Multi_GCN_Multihead_att_new as m\n"
        "assert m.__name__ == 'mgnns_b200.api.multi_gcn'\n"
        "print('ok')\n" % ROOT)
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith('ok'), out.stderr[-2000:]


WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from mgnns_b200.ddp import GradientAllReducer
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo', rank=rank, world_size=world)
torch.manual_seed(0)
model = torch.nn.Sequential(torch.nn.Linear(20, 64), torch.nn.ReLU(), torch.nn.Linear(64, 64), torch.nn.ReLU(),
                            torch.nn.Linear(64, 5))
unused = torch.nn.Linear(3, 3)                     # never receives a gradient, must stay out of the buckets
model.add_module('unused', unused)
red = GradientAllReducer(model, bucket_bytes=8 * 1024)
g = torch.Generator().manual_seed(100)
X = torch.randn(8, 20, generator=g); Y = torch.randint(0, 5, (8,), generator=g)
xs, ys = X[rank * 4:(rank + 1) * 4], Y[rank * 4:(rank + 1) * 4]
ref = torch.nn.Sequential(*[m for n, m in model.named_children() if n != 'unused'])
for step in range(3):
    model.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(model[:5](xs) if False else ref(xs), ys)
    loss.backward()
    red.finish()
    # full-batch gradient on one process == average of the two half-batch gradients
    full = [p.detach().clone().requires_grad_() for p in ref.parameters()]
    h = torch.relu(torch.nn.functional.linear(X, full[0], full[1]))
    h = torch.relu(torch.nn.functional.linear(h, full[2], full[3]))
    out = torch.nn.functional.linear(h, full[4], full[5])
    torch.nn.functional.cross_entropy(out, Y).backward()
    for p, f in zip(ref.parameters(), full):
        assert torch.allclose(p.grad, f.grad, atol=1e-6), (step, (p.grad - f.grad).abs().max())
    assert unused.weight.grad is None
    assert len(red.buckets) >= 2
    with torch.no_grad():
        for p in ref.parameters():
            p -= 0.1 * p.grad
dist.barrier()
if rank == 0:
    print('ddp ok', red.payload_bytes())
dist.destroy_process_group()
'''


def test_gradient_allreducer_world_size_2_gloo(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29531', WORLD_SIZE='2')
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-3000:]
    assert 'ddp ok' in outs[0][0]
