"""sys.modules aliases and stubs that let the reference's entry script and training engine import
unchanged against mgnns_b200 (SURVEY §2.3 lists the as-shipped defects this absorbs).

    import mgnns_b200.compat as compat
    compat.install(reference_root='/path/to/MGNNS')

Nothing here is on the compute path.
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__mgnns_stub__ = True
    sys.modules[name] = mod
    return mod


def _have(name):
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


class _AverageValueMeter:
    """Minimal torchnet.meter.AverageValueMeter (used by the engine for loss/time averages, engine:102-105)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.n, self.sum, self.var_acc = 0, 0.0, 0.0

    def add(self, value, n=1):
        self.sum += float(value) * n
        self.var_acc += float(value) ** 2 * n
        self.n += n

    def value(self):
        if self.n == 0:
            return float('nan'), float('nan')
        mean = self.sum / self.n
        std = max(self.var_acc / self.n - mean * mean, 0.0) ** 0.5
        return mean, std


def install(reference_root=None):
    """Register aliases; `reference_root` (optional) is put on sys.path so that the reference's own
    engine/dataset/utils modules — which stay the reference's — can be imported."""
    from .api import graph_util, layers, multi_gcn, pmi, text_gcn, vocab

    if not hasattr(np, 'int'):
        np.int = int            # utils/util.py:397 uses the alias numpy removed in 1.24

    # third-party packages the reference imports unconditionally but that are not installed
    if not _have('dgl'):
        fn = _stub('dgl.function')
        _stub('dgl', function=fn)
    if not _have('word2vec'):
        _stub('word2vec', load=lambda path: (_ for _ in ()).throw(FileNotFoundError(path)))
    if not _have('torchnet'):
        meter = _stub('torchnet.meter', AverageValueMeter=_AverageValueMeter)
        _stub('torchnet', meter=meter)
    if not _have('apex'):
        amp = _stub('apex.amp')
        _stub('apex', amp=amp)

    models = sys.modules.get('models') or types.ModuleType('models')
    models.__path__ = getattr(models, '__path__', [])
    sys.modules['models'] = models
    for name, mod in (('Multi_GCN_Multihead_att', multi_gcn), ('Multi_GCN_Multihead_att_new', multi_gcn),
                      ('Text_GCN', text_gcn), ('submodules', layers), ('moudles', layers)):
        sys.modules['models.' + name] = mod
        setattr(models, name, mod)
    mha_pkg = types.ModuleType('models.multi_head_att')
    mha_pkg.submodules = layers
    sys.modules['models.multi_head_att'] = mha_pkg
    sys.modules['models.multi_head_att.submodules'] = layers

    if reference_root is not None:
        reference_root = os.path.abspath(reference_root)
        if reference_root not in sys.path:
            sys.path.insert(0, reference_root)
        import utils as ref_utils                      # the reference's package (namespace or regular)
        import utils.util as ref_util
        ref_util.gen_A = graph_util.gen_A
        ref_util.gen_adj = graph_util.gen_adj
        sys.modules['utils.pmi'] = pmi
        sys.modules['utils.vocab'] = vocab
        sys.modules['utils.vocab_new'] = vocab
        ref_utils.pmi, ref_utils.vocab, ref_utils.vocab_new = pmi, vocab, vocab
        try:
            ds = importlib.import_module('utils.Multi_GCN_Co_att_dataset')
            sys.modules['utils.Multi_GCN_Co_att_dataset_new'] = ds
        except Exception:          # dataset needs PIL/word2vec data that may be absent; the model path does not
            pass
    else:
        utils = sys.modules.get('utils') or types.ModuleType('utils')
        utils.__path__ = getattr(utils, '__path__', [])
        sys.modules['utils'] = utils
        util_mod = _stub('utils.util', gen_A=graph_util.gen_A, gen_adj=graph_util.gen_adj)
        for name, mod in (('util', util_mod), ('pmi', pmi), ('vocab', vocab), ('vocab_new', vocab)):
            sys.modules['utils.' + name] = mod
            setattr(utils, name, mod)
    return True
