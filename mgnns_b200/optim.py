"""clip_grad_norm_ + Adam of the training step as two kernels over flat buffers
(ref: engine/Multi_GCN_Multihead_Att_engine.py:850-851 and the optimizer built at Tumblr_Multi_GCN_Multihead_Att.py:164
 from model.get_config_optim: torch.optim.Adam with per-group learning rates and L2 weight decay).

`FlatClipAdam` is built FROM a torch.optim.Adam instance (same parameter groups, lr, betas, eps, weight_decay), moves the
optimizer-owned parameters into one contiguous buffer (each nn.Parameter becomes a view of it; state_dict and
everything else see the same tensors) and keeps the two Adam moments in two more.  `step(flat_grads)` then runs
mgnns_sqnorm_f32 + mgnns_clip_adam_f32: the global gradient norm over ALL parameters that receive gradients (the
never-stepped ones included, as clip_grad_norm_(model.parameters()) does), the in-place scaling of every
gradient, and torch's Adam formula for the owned parameters.  Everything is device resident (step counter included), so
the step can be captured in a CUDA graph.  Tested against clip_grad_norm_ + torch.optim.Adam on the real model.
"""
import torch

from . import _abi

_lib = _abi.lib
_check = _abi.check


def _pad4(n):
    return (n + 3) // 4 * 4


class FlatGradients:
    """One flat fp32 buffer for the gradients of every parameter that receives one (in model.parameters() order, each
    tensor starting at a multiple of 4 floats); after pack() every p.grad aliases its slice, so the NCCL all-reduce,
    the clip and the optimizer all work on the same memory."""

    def __init__(self, params, alloc=None):
        """alloc(numel) -> float32 tensor of that many zeros: lets the buffer live in peer-mapped memory
        (mgnns_b200.p2p.PeerAllReduce) so that the all-reduce is a kernel over NVLink on this very buffer."""
        self.params = [p for p in params if p.grad is not None]
        if not self.params:
            raise RuntimeError("FlatGradients: run one backward pass first (parameters without a gradient are left out)")
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += _pad4(p.numel())
        dev = self.params[0].device
        self.flat = torch.zeros(off, device=dev, dtype=torch.float32) if alloc is None else alloc(off)
        if self.flat.numel() != off or self.flat.dtype != torch.float32 or self.flat.device != dev:
            raise RuntimeError("FlatGradients: alloc() must return %d float32 elements on %s" % (off, dev))
        self.peer = None            # set by the owner when the buffer is a PeerAllReduce's
        self.views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(self.offsets, self.params)]

    def pack(self):
        """Multi-tensor copy of the current .grad tensors into the buffer (a no-op for those that already alias it)."""
        src = [p.grad for p in self.params]
        todo = [(v, g) for v, g in zip(self.views, src) if g is not None and g.data_ptr() != v.data_ptr()]
        if todo:
            torch._foreach_copy_([v for v, _ in todo], [g for _, g in todo])
        for p, v in zip(self.params, self.views):
            p.grad = v


class FlatClipAdam:
    def __init__(self, optimizer: torch.optim.Adam, flat_grads: FlatGradients, max_norm: float):
        if not isinstance(optimizer, torch.optim.Adam):
            raise TypeError("FlatClipAdam mirrors torch.optim.Adam")
        self.grads = flat_grads
        self.max_norm = float(max_norm)
        groups = optimizer.param_groups
        g0 = groups[0]
        if any(g['amsgrad'] or g['maximize'] or g['betas'] != g0['betas'] or g['eps'] != g0['eps'] for g in groups):
            raise NotImplementedError("FlatClipAdam: one (betas, eps) pair, no amsgrad / maximize")
        self.beta1, self.beta2, self.eps = float(g0['betas'][0]), float(g0['betas'][1]), float(g0['eps'])
        owned = {}
        for g in groups:
            for p in g['params']:
                owned[p] = (float(g['lr']), float(g['weight_decay']))
        dev = flat_grads.flat.device
        # parameter / moment buffers: the owned parameters that receive gradients, in gradient-buffer order
        p_off, off = {}, 0
        for p in flat_grads.params:
            if p in owned:
                p_off[p] = off
                off += _pad4(p.numel())
        self.p_flat = torch.zeros(off, device=dev, dtype=torch.float32)
        self.m_flat = torch.zeros(off, device=dev, dtype=torch.float32)
        self.v_flat = torch.zeros(off, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, o in p_off.items():
                view = self.p_flat[o:o + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view                         # the nn.Parameter now lives in the flat buffer
        seg_g = [o for o in flat_grads.offsets] + [flat_grads.flat.numel()]
        seg_p = [p_off.get(p, -1) for p in flat_grads.params]
        seg_lr = [owned.get(p, (0.0, 0.0))[0] for p in flat_grads.params]
        seg_wd = [owned.get(p, (0.0, 0.0))[1] for p in flat_grads.params]
        self.n_seg = len(seg_p)
        self.seg_g = torch.tensor(seg_g, device=dev, dtype=torch.int64)
        self.seg_p = torch.tensor(seg_p, device=dev, dtype=torch.int64)
        self.seg_lr = torch.tensor(seg_lr, device=dev, dtype=torch.float32)
        self.seg_wd = torch.tensor(seg_wd, device=dev, dtype=torch.float32)
        self.sqnorm = torch.zeros(1, device=dev, dtype=torch.float64)
        self.step_count = torch.zeros(1, device=dev, dtype=torch.int64)
        self.owned_params = list(p_off)
        self.skipped = [p for p in owned if p not in p_off]      # owned but never receive a gradient: torch skips them too

    def zero_grad(self):
        """optimizer.zero_grad() of the reference (engine:841): only the optimizer-owned gradients are cleared; the
        others keep accumulating (they alias the flat buffer, autograd adds in place)."""
        for p in self.owned_params:
            p.grad = None

    def step(self):
        """Gradients must be packed (FlatGradients.pack) and, when world > 1, already all-reduced and averaged."""
        g = self.grads.flat
        s = torch.cuda.current_stream(g.device).cuda_stream
        self.step_count.add_(1)
        _check(_lib.mgnns_sqnorm_f32(g.data_ptr(), g.numel(), self.sqnorm.data_ptr(), s), "sqnorm")
        _check(_lib.mgnns_clip_adam_f32(g.data_ptr(), g.numel(), self.seg_g.data_ptr(), self.seg_p.data_ptr(),
                                        self.seg_lr.data_ptr(), self.seg_wd.data_ptr(), self.n_seg, self.p_flat.data_ptr(),
                                        self.m_flat.data_ptr(), self.v_flat.data_ptr(), self.sqnorm.data_ptr(), self.max_norm,
                                        self.beta1, self.beta2, self.eps, self.step_count.data_ptr(), s), "clip_adam")

    def total_norm(self):
        """The gradient norm the last step() clipped with (host sync)."""
        return float(self.sqnorm.sqrt().item())
