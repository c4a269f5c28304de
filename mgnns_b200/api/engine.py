"""Per-batch training / evaluation loop with the reference engine's hooks and `state` keys, metrics on the device
(ref: engine/Multi_GCN_Multihead_Att_engine.py — Engine :23-312, train/validate/test :511-652,
 MultiClassEngine :655-788, GCNMultiClassEngine :791-865).

The reference's `on_forward` copies the predictions to the host and calls sklearn four times EVERY batch
(engine:829-838: `.cpu()` sync + accuracy_score + three f1_score), and `on_end_batch` calls `loss.item()`
(engine:183) — three host round trips per batch that leave the GPU idle.  Here `on_forward` makes the same model
call (same seven positional arguments, engine:825) and the same optimizer sequence (engine:840-851), but the
arg-max and a per-batch confusion matrix are computed by `mgnns_confusion_count` into slot `iteration` of a device
buffer, the loss stays a device scalar, and everything the reference exposes through `state` — `batch_acc_list`,
`batch_micro_f1_list`, `batch_macro_f1_list`, `batch_weighted_f1_list`, `meter_loss`, `id_list`, `pred_list`,
`target_list`, `epoch_acc`, ... — is filled at `on_end_epoch` from ONE device->host copy, with the same values
(accuracy / F1 are functions of the confusion matrix; `metrics_from_confusion` restates sklearn's definitions and is
tested against sklearn).  Per-batch display (`print_freq`) still works: it flushes what has accumulated so far.

Out of scope (callers / host I/O): `learning()` with its DataLoaders, transforms, checkpoints and result
files — use the reference's engine for those and pass this class's hooks, or call train()/validate() directly.
"""
import time

import numpy as np
import torch
import torch.nn as nn

from .. import ops


def metrics_from_confusion(conf):
    """(accuracy, micro-F1, macro-F1, weighted-F1) of one batch from its confusion matrix conf[target, pred].

    Same definitions as the sklearn calls at engine:835-838: the label set is the union of the labels present in
    the targets and in the predictions; per-label F1 is 0 when precision + recall is 0; 'macro' is the plain mean
    over that set, 'weighted' weighs by support (targets per label), 'micro' equals accuracy for single-label
    multi-class data."""
    conf = np.asarray(conf, dtype=np.float64)
    n = conf.sum()
    if n == 0:
        return float('nan'), float('nan'), float('nan'), float('nan')
    tp = np.diag(conf)
    support, predicted = conf.sum(1), conf.sum(0)
    present = (support + predicted) > 0
    denom = support + predicted
    f1 = np.where(denom > 0, 2.0 * tp / np.where(denom > 0, denom, 1.0), 0.0)
    acc = tp.sum() / n
    macro = f1[present].mean()
    weighted = (f1 * support).sum() / support.sum() if support.sum() > 0 else 0.0
    return float(acc), float(acc), float(macro), float(weighted)


class _Meter:
    """torchnet.meter.AverageValueMeter's value()/add()/reset() (engine:102-105)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.n, self.sum, self.sq = 0, 0.0, 0.0

    def add(self, value, n=1):
        self.sum += float(value) * n
        self.sq += float(value) ** 2 * n
        self.n += n

    def value(self):
        if self.n == 0:
            return float('nan'), float('nan')
        mean = self.sum / self.n
        return mean, max(self.sq / self.n - mean * mean, 0.0) ** 0.5


class GCNMultiClassEngine:
    """Drop-in for the reference's GCNMultiClassEngine hooks (engine:791-865) with device-side metrics."""

    _DEFAULTS = dict(image_size=224, batch_size=16, workers=25, device_ids=[1], evaluate=False, start_epoch=0,
                     max_epochs=90, object_t_value=0.4, place_t_value=0.4, epoch_acc=0, epoch_micro_f1=0,
                     epoch_macro_f1=0, epoch_weighted_f1=0, fp16=False, fp16_opt_level='O1', use_pb=False,
                     print_freq=0, max_batches_per_epoch=4096)

    def __init__(self, state=None):
        self.state = {} if state is None else state
        if self._state('use_gpu') is None:
            self.state['use_gpu'] = torch.cuda.is_available()
        for k, v in self._DEFAULTS.items():
            if self._state(k) is None:
                self.state[k] = v
        for k in ('epoch_step', 'batch_acc_list', 'batch_micro_f1_list', 'batch_macro_f1_list', 'batch_weighted_f1_list',
                  'id_list', 'pred_list', 'target_list'):
            if self._state(k) is None:
                self.state[k] = []
        if self.state['fp16']:
            raise NotImplementedError("mgnns_b200 engine: apex AMP (fp16=True) is outside the fp32 hot path")
        self.state['meter_loss'] = _Meter()
        self.state['batch_time'] = _Meter()
        self.state['data_time'] = _Meter()
        self._conf = None                  # int32 [max_batches, C, C] on the device
        self._pending = []                 # per batch since the last flush: (slot, loss, pred, target, ids)
        self._flushed = 0

    def _state(self, name):
        return self.state.get(name)

    # ------------------------------------------------------------------ epoch hooks (engine:116-128, :130-170)
    def on_start_epoch(self, training, model, criterion, data_loader, optimizer=None, display=True):
        for k in ('meter_loss', 'batch_time', 'data_time'):
            self.state[k].reset()
        for k in ('batch_acc_list', 'batch_micro_f1_list', 'batch_macro_f1_list', 'batch_weighted_f1_list',
                  'id_list', 'target_list', 'pred_list'):
            self.state[k].clear()
        self._pending = []
        self._flushed = 0
        if self._conf is not None:
            self._conf.zero_()

    def flush_metrics(self):
        """One device->host copy for every batch accumulated since the last flush; fills the reference's lists."""
        if not self._pending:
            return
        slots = [p[0] for p in self._pending]
        conf = self._conf[slots[0]:slots[-1] + 1].cpu().numpy()              # the single sync
        losses = torch.stack([p[1] for p in self._pending]).cpu().tolist()
        preds = torch.cat([p[2] for p in self._pending]).cpu().numpy()
        targets = torch.cat([p[3] for p in self._pending]).cpu().numpy()
        off = 0
        for i, (slot, _, pred, _, ids) in enumerate(self._pending):
            acc, micro, macro, weighted = metrics_from_confusion(conf[slot - slots[0]])
            self.state['loss_batch'] = losses[i]
            self.state['meter_loss'].add(losses[i])
            for key, v in (('acc', acc), ('micro_f1', micro), ('macro_f1', macro), ('weighted_f1', weighted)):
                self.state['batch_' + key] = v
                self.state['batch_%s_list' % key].append(v)
            n = pred.numel()
            self.state['id_list'] = self.state['id_list'] + list(ids)
            self.state['pred_list'] = self.state['pred_list'] + preds[off:off + n].tolist()
            self.state['target_list'] = self.state['target_list'] + targets[off:off + n].tolist()
            off += n
        self.state['acc'], self.state['micro_f1'] = self.state['batch_acc'], self.state['batch_micro_f1']
        self.state['macro_f1'], self.state['weighted_f1'] = self.state['batch_macro_f1'], self.state['batch_weighted_f1']
        self.state['pred'] = preds[off - n:off]
        self._flushed += len(self._pending)
        self._pending = []

    def on_end_epoch(self, training, model, criterion, data_loader, optimizer=None, display=True):
        self.flush_metrics()
        loss = self.state['meter_loss'].value()[0]
        n = len(data_loader)
        for key in ('acc', 'micro_f1', 'macro_f1', 'weighted_f1'):
            self.state['epoch_' + key] = sum(self.state['batch_%s_list' % key]) / n
        acc, micro_f1 = self.state['epoch_acc'], self.state['epoch_micro_f1']
        macro_f1, weighted_f1 = self.state['epoch_macro_f1'], self.state['epoch_weighted_f1']
        if display:
            head = '-----------------Epoch: [%s]' % self.state.get('epoch') if training else '--------------------Val: '
            print('%s\tLoss %.4f\tAcc %.4f\tMicro_f1 %.4f\tMacro_f1 %.4f\tWeighted_f1 %.4f'
                  % (head, loss, acc, micro_f1, macro_f1, weighted_f1))
        return (loss, acc, micro_f1, macro_f1, weighted_f1, self.state['id_list'], self.state['target_list'],
                self.state['pred_list'])

    # ------------------------------------------------------------------ batch hooks
    def on_start_batch(self, training, model, criterion, data_loader, optimizer=None, display=True):
        """Unpack the dataset tuple (engine:853-865; the same image feeds both trunks, :861-862)."""
        inp = self.state['input']
        self.state['id'] = inp[0]
        self.state['text_feature'] = inp[2]
        self.state['text_lens'] = inp[3]
        self.state['text_mask'] = inp[4]
        self.state['object_feature'] = inp[5]
        self.state['place_feature'] = inp[5]
        self.state['image_name'] = inp[6]
        self.state['object_input'] = inp[7]
        self.state['place_input'] = inp[8]

    def model_args(self, device):
        """The seven positional arguments of the model call, moved and cast as engine:793-811 does.  text_lens stays
        on the host: the reference moves it to the device (engine:805) only for pack_padded_sequence to copy it back
        (model:376), a sync the LSTM plan does not need."""
        s = self.state
        nb = s['use_gpu']
        return (s['text_feature'].to(device, non_blocking=nb), s['text_lens'], s['text_mask'].to(device, non_blocking=nb),
                s['object_feature'].float().to(device, non_blocking=nb), s['place_feature'].float().to(device, non_blocking=nb),
                s['object_input'].float().detach().to(device, non_blocking=nb),
                s['place_input'].float().detach().to(device, non_blocking=nb))

    def on_forward(self, training, model, criterion, data_loader, optimizer=None, display=True):
        device = torch.device('cuda:0' if self.state['use_gpu'] else 'cpu')
        if device.type != 'cuda':
            raise RuntimeError("mgnns_b200 engine: the hot path is CUDA only (no CPU fallback)")
        args = self.model_args(device)
        target = self.state['target'].to(device).long()
        with torch.set_grad_enabled(training):
            logits = model(*args)                                            # engine:825
            self.state['loss'] = criterion(logits, target)                   # engine:826
            self.state['output'] = torch.nn.functional.softmax(logits, dim=1)   # engine:828
        scores = self.state['output'].detach()
        B, C = scores.shape
        slot = int(self.state.get('iteration', len(self._pending) + self._flushed))
        if self._conf is None or self._conf.shape[1] != C or slot >= self._conf.shape[0]:
            self.flush_metrics()
            cap = max(self.state['max_batches_per_epoch'], slot + 1)
            self._conf = torch.zeros((cap, C, C), device=device, dtype=torch.int32)
        pred = torch.empty((B,), device=device, dtype=torch.int64)
        ops.confusion_count(scores, target, self._conf[slot], pred)          # replaces engine:830-838
        self.state['pred_device'] = pred
        self._pending.append((slot, self.state['loss'].detach(), pred, target, self.state['id']))
        if training:
            optimizer.zero_grad()                                            # engine:841 (set_to_none left to torch's default)
            self.state['loss'].backward()
            reducer = self.state.get('grad_reducer')
            if reducer is not None:
                reducer.finish()                                             # the one collective, before the clip
            nn.utils.clip_grad_norm_(model.parameters(), max_norm=10.0)      # engine:850
            optimizer.step()

    def on_end_batch(self, training, model, criterion, data_loader, optimizer=None, display=True):
        """engine:180-200 without the host round trips; the reference's per-batch print is kept behind print_freq."""
        pf = self.state['print_freq']
        if display and pf and self.state.get('iteration', 0) % pf == 0:
            self.flush_metrics()
            loss = self.state['meter_loss'].value()[0]
            print('%s: [%s/%d]\tLoss %.4f (%.4f)\tAcc %.4f\tMicro_f1 %.4f\tMacro_f1 %.4f\tWeighted_f1 %.4f'
                  % ('Epoch [%s]' % self.state.get('epoch') if training else 'Val', self.state.get('iteration'),
                     len(data_loader), self.state['loss_batch'], loss, self.state['batch_acc'], self.state['batch_micro_f1'],
                     self.state['batch_macro_f1'], self.state['batch_weighted_f1']))

    # ------------------------------------------------------------------ loops (engine:511-652)
    def _run(self, training, data_loader, model, criterion, optimizer=None):
        model.train() if training else model.eval()
        self.on_start_epoch(training, model, criterion, data_loader, optimizer)
        end = time.time()
        for i, (inp, target) in enumerate(data_loader):
            self.state['iteration'] = i
            self.state['data_time_batch'] = time.time() - end
            self.state['data_time'].add(self.state['data_time_batch'])
            self.state['input'] = inp
            self.state['target'] = target
            self.on_start_batch(training, model, criterion, data_loader, optimizer)
            if self.state['use_gpu']:
                self.state['target'] = self.state['target'].cuda(non_blocking=True)
            self.on_forward(training, model, criterion, data_loader, optimizer)
            self.state['batch_time_current'] = time.time() - end
            self.state['batch_time'].add(self.state['batch_time_current'])
            end = time.time()
            self.on_end_batch(training, model, criterion, data_loader, optimizer)
        return self.on_end_epoch(training, model, criterion, data_loader, optimizer, display=bool(self.state['print_freq']))

    def train(self, data_loader, model, criterion, optimizer, epoch):
        self.state['epoch'] = epoch
        return self._run(True, data_loader, model, criterion, optimizer)

    def validate(self, data_loader, model, criterion):
        return self._run(False, data_loader, model, criterion)[:5]

    def test(self, data_loader, model, criterion):
        return self._run(False, data_loader, model, criterion)


# ---------------------------------------------------------------------- label-graph adjacency from data
def _encode_label_lists(objects):
    n = len(objects)
    lens = np.array([len(o) for o in objects], dtype=np.int32)
    width = int(lens.max()) if n else 0
    arr = np.full((n, max(width, 1)), -1, dtype=np.int32)
    for i, o in enumerate(objects):
        if len(o):
            arr[i, :len(o)] = np.asarray(list(o), dtype=np.int64)
    return arr, lens


def label_counts(objects, num_classes, device=None):
    """(nums float64 [C], Adj float64 [C,C]) of a list of per-image label lists, counted on the GPU
    (ref: utils/util.py:336-357 generate_nums + generate_Adj, one pass instead of two Python loops)."""
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device())
    arr, lens = _encode_label_lists(objects)
    nums = torch.zeros((num_classes,), device=device, dtype=torch.int64)
    adj = torch.zeros((num_classes, num_classes), device=device, dtype=torch.int64)
    if arr.shape[0]:
        ops.label_cooccurrence(torch.from_numpy(arr).to(device), torch.from_numpy(lens).to(device), num_classes, nums, adj)
    return nums.cpu().numpy().astype(np.float64), adj.cpu().numpy().astype(np.float64)


def generate_nums(objects, num_classes):
    """ref: utils/util.py:336-346"""
    return label_counts(objects, num_classes)[0]


def generate_Adj(objects, num_classes):
    """ref: utils/util.py:347-356"""
    return label_counts(objects, num_classes)[1]


def get_Adj_from_lists(splits, num_classes):
    """ref: utils/util.py:359-380 without the file I/O: `splits` is a list of per-split label-list lists; returns
    (all_nums with zeros replaced by 1, all_Adj) — the dict the reference pickles is {'nums': ..., 'adj': ...}."""
    all_nums = np.zeros(num_classes)
    all_adj = np.zeros((num_classes, num_classes))
    for objects in splits:
        nums, adj = label_counts(objects, num_classes)
        all_nums += nums
        all_adj += adj
    all_nums[all_nums == 0] = 1
    return all_nums, all_adj
