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


@pytest.mark.skipif(not os.path.isdir('/root/reference/engine'), reason="reference checkout not present")
def test_engine_hooks_feed_the_model_like_the_reference_engine():
    """The reference's own GCNMultiClassEngine (imported unchanged through compat) and mgnns_b200.api.engine are
    given the same Tumblr-shaped batch and a recording stub model on the CPU: both must hand the model the same seven
    positional arguments (engine:825) and leave the same keys in `state` (engine:853-865, :826-838)."""
    code = r"""
import sys; sys.path.insert(0, %r)
import mgnns_b200.compat as c; c.install(reference_root='/root/reference')
import importlib, torch, numpy as np
ref_eng = importlib.import_module('engine.Multi_GCN_Multihead_Att_engine')
from mgnns_b200.api.engine import GCNMultiClassEngine as Ours
from mgnns_b200 import synth
B, V = 6, 50
text, lens, mask = synth.make_texts(B, V, 100, seed=1)
img = torch.randn(B, 3, 8, 8)
oinp, pinp = synth.label_inputs(B)
inp = (['id%%d' %% i for i in range(B)], ['t'] * B, text, lens, mask, img, ['p'] * B, oinp.contiguous(), pinp.contiguous())
target = synth.make_labels(B, 7, seed=1)
calls = []
class Stub(torch.nn.Module):
    def __init__(self):
        super().__init__(); self.w = torch.nn.Parameter(torch.zeros(7))
    def forward(self, *args):
        calls.append(args)
        g = torch.Generator().manual_seed(0)
        return torch.randn(args[0].shape[0], 7, generator=g) + self.w
model, crit = Stub(), torch.nn.CrossEntropyLoss()
r = ref_eng.GCNMultiClassEngine({'use_gpu': False, 'fp16': False})
r.state['input'], r.state['target'] = inp, target
r.on_start_batch(False, model, crit, None)
r.on_forward(False, model, crit, None)
o = Ours({'use_gpu': False})
o.state['input'], o.state['target'] = inp, target
o.on_start_batch(False, model, crit, None)
ours_args = o.model_args(torch.device('cpu'))
ref_args = calls[0]
assert len(ref_args) == len(ours_args) == 7
for i, (a, b) in enumerate(zip(ref_args, ours_args)):
    assert torch.equal(a.cpu(), b.cpu()) and (a.dtype == b.dtype), i
for k in ('id', 'text_feature', 'text_lens', 'text_mask', 'object_feature', 'place_feature', 'image_name',
          'object_input', 'place_input'):
    assert (r.state[k] is o.state[k]) or r.state[k] == o.state[k], k
assert r.state['object_feature'] is r.state['place_feature']          # the same image feeds both trunks
# the reference's per-batch metrics equal ours computed from the confusion matrix of the same predictions
from mgnns_b200.api.engine import metrics_from_confusion
conf = np.zeros((7, 7)); np.add.at(conf, (target.numpy(), r.state['pred']), 1)
acc, micro, macro, weighted = metrics_from_confusion(conf)
assert abs(acc - r.state['acc']) < 1e-12 and abs(micro - r.state['micro_f1']) < 1e-12
assert abs(macro - r.state['macro_f1']) < 1e-12 and abs(weighted - r.state['weighted_f1']) < 1e-12
try:
    o.on_forward(False, model, crit, None)
except RuntimeError as e:
    assert 'CUDA only' in str(e)
else:
    raise AssertionError('CPU on_forward must fail loudly')
print('ok')
""" % ROOT
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith('ok'), (out.stdout[-500:], out.stderr[-3000:])


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
