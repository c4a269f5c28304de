"""Whole-step CUDA-graph capture of the MGNNS training step (forward + loss + backward +
gradient all-reduce + clip_grad_norm_ + optimizer step).

The eager step issues ~3,000 small launches and is bound by Python/launch overhead; replaying a
captured graph removes that.  What makes the step capturable:
  * inputs live in static device tensors (the caller copies / H2D-transfers a batch into them);
  * the LSTM schedule is a fixed-capacity LstmPlan refreshed in place (ops.LstmPlan.update_);
  * dropout seeds come from a device-resident counter advanced inside the graph (ops.set_seed_tensor);
  * the optimizer is constructed with capturable=True.
"""
import os

import torch
import torch.distributed as dist

from . import ops

# world > 1: 1 (default) = the gradient all-reduce is mgnns_allreduce_p2p_f32, one kernel over NVLink peer memory INSIDE
# the captured graph (one graph launch per step); 0 = NCCL all-reduce issued eagerly between two captured halves
_P2P_ALLREDUCE = os.environ.get("MGNNS_P2P_ALLREDUCE", "1") == "1"


class GraphedTrainStep:
    def __init__(self, model, optimizer, criterion, batch, clip_norm=10.0, world_size=1, plan_capacity=None,
                 warmup=3, flat_optimizer=None):
        """`flat_optimizer`: None = clip_grad_norm_ + optimizer.step() as the reference calls them (engine:850-851);
        True = the same arithmetic as two kernels over flat buffers (mgnns_b200.optim.FlatClipAdam, built from
        `optimizer`, which must be torch.optim.Adam); or an existing FlatClipAdam shared between several
        GraphedTrainStep objects of the same model (one per static batch)."""
        """batch: dict of STATIC tensors — text i64 [B,L], mask f32 [B,L], fo/fp f32 [B,2048,14,14],
        oinp/pinp, labels (all on the model's device) and lens (int64, CPU)."""
        self.model, self.opt, self.crit, self.batch = model, optimizer, criterion, batch
        self.clip_norm, self.world = clip_norm, world_size
        dev = batch['text'].device
        L = batch['text'].shape[1]
        n = int(batch['lens'].clamp(max=L).sum())
        # Default capacity: this batch's token count with 12.5 % headroom, rounded up to 1024 rows (the LSTM's
        # many-row products run over `capacity` rows, so B*L — the safe upper bound — would cost 6x on TumEmo-shaped
        # text); a later batch with more tokens makes update_lengths() grow the plan and re-capture (see there).
        # (world > 1: 25 % headroom — growing the plan re-captures the step, which runs collectives, so it cannot happen
        # on one rank alone; update_lengths() raises instead, see there)
        room = n // 8 if world_size == 1 else n // 4
        cap = plan_capacity if plan_capacity is not None else min(batch['text'].shape[0] * L, ((n + room + 1023) // 1024) * 1024)
        self.lens_key = batch['lens']                      # identity of this tensor keys the plan cache
        self.plan = model.make_text_plan(self.lens_key, L, capacity=cap)
        self.seed = torch.zeros(1, device=dev, dtype=torch.int64)
        self.loss = None
        self.graph = None
        self.graph_update = None
        self._flat = None
        self._views = None
        self._fg = None
        self.flat_opt = flat_optimizer
        self._warmup = warmup
        self.recaptures = 0
        self.time_allreduce = False
        self.allreduce_events = []
        self.use_p2p = world_size > 1 and _P2P_ALLREDUCE
        self._capture(warmup)

    def _capture(self, warmup):
        dev = self.batch['text'].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        prev = ops.set_seed_tensor(self.seed)
        timers = ops.KernelTimers.enabled
        ops.KernelTimers.enabled = set()
        try:
            self.graph = torch.cuda.CUDAGraph()
            if self.world == 1 or self.use_p2p:
                # one graph: with the peer-memory all-reduce the collective is just another kernel node
                with torch.cuda.graph(self.graph):
                    self.loss = self._step()
                    self.seed.add_(0x9E3779B97F4A7C15 & 0x7FFFFFFFFFFFFFFF)
            else:
                # Two captured halves with the NCCL all-reduce issued eagerly between them: the collective is then
                # an ordinary stream-ordered call of the process group (its own stream, watchdog and error
                # handling untouched by capture), at the cost of one extra graph launch per step.
                with torch.cuda.graph(self.graph):
                    self.loss = self._forward_backward()
                    self._pack_grads()
                    self.seed.add_(0x9E3779B97F4A7C15 & 0x7FFFFFFFFFFFFFFF)
                self.graph_update = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_update, pool=self.graph.pool()):
                    self._flat.div_(self.world)
                    self._clip_and_step()
        finally:
            ops.set_seed_tensor(prev)
            ops.KernelTimers.enabled = timers

    def _pack_grads(self):
        """Gather every gradient into one flat buffer with a multi-tensor copy; afterwards each p.grad aliases
        its slice (no copy back), which is what the all-reduce, clip_grad_norm_ and the optimizer then use."""
        from .optim import FlatGradients
        if self._fg is None:
            fo = self.flat_opt if (self.flat_opt is not None and self.flat_opt is not True) else None
            if fo is not None:
                self._fg = fo.grads
                self.use_p2p = self.use_p2p and self._fg.peer is not None
            elif self.use_p2p:
                from .p2p import PeerAllReduce, PeerMemoryUnavailable
                dev = self.batch['text'].device
                made = []

                def alloc(n):
                    made.append(PeerAllReduce(n, dev))       # rendezvous: every rank gets here in its first step
                    return made[0].flat
                try:
                    self._fg = FlatGradients(self.model.parameters(), alloc=alloc)
                    self._fg.peer = made[0]
                except PeerMemoryUnavailable as exc:         # raised on every rank alike: all fall back to NCCL together
                    import warnings
                    warnings.warn("mgnns_b200: peer-memory all-reduce unavailable (%s); using the NCCL all-reduce between "
                                  "two captured halves" % exc)
                    self.use_p2p = False
                    self._fg = FlatGradients(self.model.parameters())
            else:
                self._fg = FlatGradients(self.model.parameters())
            self._flat, self._views = self._fg.flat, self._fg.views
        self._fg.pack()

    def _allreduce(self):
        """The one collective of the step: all-reduce of the flat gradient buffer — the peer-memory kernel (mean, in
        place) or NCCL (SUM; the / world follows)."""
        if self.use_p2p:
            if self._fg.peer is None:
                raise RuntimeError("GraphedTrainStep: the shared FlatGradients buffer is not peer-mapped")
            self._fg.peer.all_reduce_(1.0 / self.world)
            return
        dist.all_reduce(self._flat, op=dist.ReduceOp.SUM)

    def _forward_backward(self):
        b = self.batch
        self.opt.zero_grad(set_to_none=True)
        logits = self.model(b['text'], self.lens_key, b['mask'], b['fo'], b['fp'], b['oinp'], b['pinp'])
        loss = self.crit(logits, b['labels'])
        prev = ops.defer_weight_grads(True)      # many-row weight gradients leave the critical path (ops.py)
        try:
            loss.backward()
        finally:
            ops.defer_weight_grads(prev)
            ops.join_deferred()                  # ... and are joined before anything reads .grad
        return loss.detach()

    def _clip_and_step(self):
        if self.flat_opt is None:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=self.clip_norm)
            self.opt.step()
            return
        if self.flat_opt is True:
            from .optim import FlatClipAdam
            self.flat_opt = FlatClipAdam(self.opt, self._fg, self.clip_norm)
        self.flat_opt.step()

    def _step(self):
        loss = self._forward_backward()
        if self.world > 1 or self.flat_opt is not None:
            self._pack_grads()
        if self.world > 1:
            self._allreduce()
            if not self.use_p2p:
                self._flat.div_(self.world)
        self._clip_and_step()
        return loss

    def update_lengths(self, lens_cpu):
        """New batch in the static buffers: refresh the LSTM schedule (tiny H2D on the CURRENT stream; the next
        replay() must be ordered after that stream — LstmPlan.update_ documents the staging-buffer contract).
        A batch with more valid tokens than the plan's capacity grows the plan by 1.5x and re-captures the step
        (one-off cost of a few eager steps; `recaptures` counts them) instead of failing mid-training."""
        L = self.batch['text'].shape[1]
        n = int(lens_cpu.clamp(max=L).sum())
        if n > self.plan.capacity and self.world > 1:
            # Re-capturing runs warm-up steps with the gradient all-reduce in them; a rank that did so alone would
            # leave the others waiting in a collective (NCCL) or trip the 20 s barrier time-out of the peer-memory
            # kernel.  Fail loudly instead: the caller picks a capacity every rank can live with.
            raise RuntimeError("GraphedTrainStep.update_lengths: %d valid tokens exceed the plan capacity %d on this rank; "
                               "with world_size > 1 the step cannot be re-captured by one rank alone — construct it with "
                               "plan_capacity=batch*max_len (always safe) or a bound that holds for every batch"
                               % (n, self.plan.capacity))
        if n > self.plan.capacity:
            cap = min(self.batch['text'].shape[0] * L, ((max(n, self.plan.capacity * 3 // 2) + 1023) // 1024) * 1024)
            torch.cuda.synchronize()
            self.lens_key.copy_(lens_cpu)                  # in-place: bumps the version, so the plan cache misses on purpose
            self.plan = self.model.make_text_plan(self.lens_key, L, capacity=cap)
            if self.flat_opt is None and not self.use_p2p:      # (a peer-mapped buffer is kept: making one is a collective)
                self._flat, self._views, self._fg = None, None, None
                for p in self.model.parameters():
                    p.grad = None
            self.recaptures += 1
            self._capture(max(1, self._warmup - 1))
            return
        self.plan.update_(lens_cpu)

    def replay(self):
        self.graph.replay()
        if self.world > 1 and not self.use_p2p:
            if self.time_allreduce:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                self._allreduce()
                e1.record()
                self.allreduce_events.append((e0, e1))
            else:
                self._allreduce()
            self.graph_update.replay()
        return self.loss
