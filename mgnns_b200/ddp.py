"""Batch-sharded data parallelism: one process per GPU, one collective per training step — the
gradient all-reduce (SUM then / world), bucketed and overlapped with backward.

The reference has no distributed code (nn.DataParallel is commented out, engine:365); this is the
new piece that sits between loss.backward() and clip_grad_norm_ (engine:846-850).  Parameters that
never receive a gradient (8.3 M of them) are discovered on the first step and left out
of the buckets; parameters that get gradients but are never stepped are reduced like the others,
because clip_grad_norm_ folds their norm into the clip coefficient.
"""
import torch
import torch.distributed as dist


class GradientAllReducer:
    def __init__(self, model, process_group=None, bucket_bytes=32 << 20):
        self.model = model
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.bucket_bytes = bucket_bytes
        self.buckets = None          # list of dict(flat, params, views, pending, work)
        self.index = {}              # param -> (bucket id, slot)
        self.hooks = []

    # -- first step: learn which parameters receive gradients, build buckets in reverse order ------
    def _build(self):
        params = [p for p in self.model.parameters() if p.grad is not None]
        params.reverse()             # backward produces gradients roughly in reverse registration order
        self.buckets = []
        cur, cur_bytes = [], 0
        for p in params:
            nbytes = p.numel() * p.element_size()
            if cur and cur_bytes + nbytes > self.bucket_bytes:
                self._close(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self._close(cur)
        for p in params:
            self.hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _close(self, params):
        total = sum(p.numel() for p in params)
        flat = torch.zeros(total, device=params[0].device, dtype=params[0].dtype)
        views, off = [], 0
        for p in params:
            views.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        bid = len(self.buckets)
        for slot, p in enumerate(params):
            self.index[p] = (bid, slot)
        self.buckets.append(dict(flat=flat, params=params, views=views, pending=len(params), work=None,
                                 events=[None] * len(params)))

    def _on_grad(self, p):
        """Post-accumulate-grad hook.  With model.branch_streams the AccumulateGrad nodes of one bucket run on
        different CUDA streams (the stream of each parameter's forward op), so every copy into the bucket records an
        event on ITS stream, and the all-reduce — issued from whichever hook happens to complete the bucket — first
        makes its launching stream wait on all of them; otherwise NCCL could read slots that are still being written."""
        bid, slot = self.index[p]
        b = self.buckets[bid]
        b['views'][slot].copy_(p.grad)
        if p.is_cuda:
            ev = b['events'][slot]
            if ev is None:
                ev = b['events'][slot] = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(p.device))
        b['pending'] -= 1
        if b['pending'] == 0 and self.world > 1:
            self._launch(b)

    def _launch(self, b):
        if b['flat'].is_cuda:
            cur = torch.cuda.current_stream(b['flat'].device)
            for ev in b['events']:
                if ev is not None:
                    cur.wait_event(ev)
        b['work'] = dist.all_reduce(b['flat'], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    # -- call after loss.backward(), before clip_grad_norm_ ------------------------------------------
    def finish(self):
        if self.buckets is None:
            self._build()
            for b in self.buckets:           # first step: reduce synchronously
                for v, p in zip(b['views'], b['params']):
                    v.copy_(p.grad)
                if self.world > 1:
                    dist.all_reduce(b['flat'], op=dist.ReduceOp.SUM, group=self.group)
        else:
            for b in self.buckets:
                if b['pending'] != 0:
                    # a parameter produced no gradient this step: reduce what is there
                    if self.world > 1:
                        self._launch(b)
                if b['work'] is not None:
                    b['work'].wait()
                    b['work'] = None
        for b in self.buckets:
            if b['flat'].is_cuda:
                cur = torch.cuda.current_stream(b['flat'].device)
                for ev in b['events']:
                    if ev is not None:
                        cur.wait_event(ev)
            if self.world > 1:
                b['flat'].div_(self.world)
            for v, p in zip(b['views'], b['params']):
                p.grad = v                    # averaged gradient, aliasing the bucket (no copy back)
            b['pending'] = len(b['params'])

    def payload_bytes(self):
        return 0 if self.buckets is None else sum(b['flat'].numel() * 4 for b in self.buckets)

    def remove(self):
        for h in self.hooks:
            h.remove()
        self.hooks = []
