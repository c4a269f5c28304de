"""torch.library ops ("mgnns::*") over the C-ABI, with autograd formulas.

Every op here:
  * checks device/dtype/contiguity up front and raises on CPU tensors (there is no
    CPU fallback and no other backend),
  * allocates its outputs through torch's caching allocator,
  * enqueues the hand-written kernels on torch's current CUDA stream,
  * has a fake (meta) implementation, and a backward that is itself made of
    C-ABI calls.
"""
import math
import os
import re
from typing import List, Optional, Tuple

import torch

from . import _abi

_lib = _abi.lib
_check = _abi.check

ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2

# Precision mode of the two streaming image-bank contractions (everything else is fp32 FMA):
#   "fp32"   exact fp32 CUDA-core GEMM (parity reference)
#   "tf32x3" tcgen05 tensor cores with the 3xTF32 operand split (fp32-class accuracy)  [default]
#   "tf32"   tcgen05 tensor cores, plain TF32 operands (looser bound)
_PRECISIONS = {"fp32": None, "tf32x3": 1, "tf32": 0}
_precision = os.environ.get("MGNNS_PRECISION", "tf32x3")
if _precision not in _PRECISIONS:
    raise RuntimeError("MGNNS_PRECISION must be one of %s" % sorted(_PRECISIONS))


def set_precision(mode: str) -> str:
    """Select the contraction mode; returns the previous one."""
    global _precision
    if mode not in _PRECISIONS:
        raise ValueError("precision must be one of %s" % sorted(_PRECISIONS))
    prev, _precision = _precision, mode
    return prev


def get_precision() -> str:
    return _precision


def _tc_ok(C, P, O, o_max):
    return _PRECISIONS[_precision] is not None and C % 32 == 0 and P % 4 == 0 and O % 4 == 0 and O <= o_max

_LIB = torch.library.Library("mgnns", "DEF")


# ----------------------------------------------------------------------------- kernel timers
class KernelTimers:
    """Optional CUDA-event brackets around named kernel launches (bench.py's live roofline numbers).
    Disabled by default: no events are recorded unless a name is enabled."""
    enabled = set()
    records = {}

    @classmethod
    def reset(cls, names=()):
        cls.enabled = set(names)
        cls.records = {n: [] for n in names}

    @classmethod
    def mean_ms(cls, name):
        ev = cls.records.get(name, [])
        return (sum(a.elapsed_time(b) for a, b in ev) / len(ev), len(ev)) if ev else (None, 0)


class _timed:
    __slots__ = ("name", "start")

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if self.name in KernelTimers.enabled:
            self.start = torch.cuda.Event(enable_timing=True)
            self.start.record()
        else:
            self.start = None

    def __exit__(self, *exc):
        if self.start is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            KernelTimers.records[self.name].append((self.start, end))
        return False


# ----------------------------------------------------------------------------- deferred weight gradients
# The gradient of a weight is not needed before the optimizer, but autograd computes it inline, in the middle of
# the chain that produces the gradient the NEXT layer is waiting for (LSTM layer 2 -> layer 1: three many-row
# products, ~0.5 ms per layer on the critical path).  With deferral on, the LSTM weight-gradient products are
# enqueued on a side stream (they overlap the latency-bound recurrence of the next layer), autograd is handed None
# for those parameters, and join_deferred() — called after loss.backward() — joins the side stream and adds the
# results to .grad.  Opt-in (GraphedTrainStep does both); gradients are identical to the inline path.
#
# "Heavy" deferred work (the image-bank weight gradients: persistent tensor-core kernels that take every SM for
# ~0.8 ms) is additionally GATED: the latency-bound LSTM recurrence is the critical path of the backward pass, and a
# persistent kernel that got onto the SMs first made it wait for the whole 0.8 ms.  The LSTM backward therefore drops
# a gate event just before each recurrence launch, heavy jobs are only queued while autograd runs, and
# join_deferred() issues job i on its own side stream behind (its inputs, gate i + a few microseconds): the
# recurrence has been handed its SMs when the heavy kernel starts, which then fills the remaining ones.  Everything
# is stream dependencies, so a CUDA-graph capture records exactly this order.
_defer_state = {"enabled": False, "streams": {}, "pending": [], "heavy": [], "gates": [], "gate_streams": [], "before_heavy": [], "lazy": []}
_HEAVY_AFTER_TEXT = os.environ.get("MGNNS_HEAVY_AFTER_TEXT", "0") == "1"   # measured: 7.44 vs 7.36 ms without
_GATE_DELAY_NS = int(os.environ.get("MGNNS_GATE_DELAY_NS", "30000"))
_DEFER_SMALL = os.environ.get("MGNNS_DEFER_SMALL", "1") == "1"     # also defer the small (M = batch) weight gradients


def _parse_heavy_plan(text):
    """"g0,g1" -> [(0 CTAs = full grid, gate 0), (0, gate 1)]; "c100,g0" -> job 0 at once on <= 100 CTAs, job 1 behind
    gate 0; "off" -> heavy jobs run inline.  Jobs beyond the list reuse its last entry."""
    if text.strip() == "off":
        return None
    plan = []
    for spec in text.split(","):
        m = re.fullmatch(r"(?:c(\d+))?(?:g(\d+))?", spec.strip())
        if m is None:
            raise ValueError("MGNNS_HEAVY_PLAN: bad entry %r" % spec)
        plan.append((int(m.group(1) or 0), int(m.group(2)) if m.group(2) is not None else -1))
    return plan


_HEAVY_PLAN = _parse_heavy_plan(os.environ.get("MGNNS_HEAVY_PLAN", "g0,g1"))


def defer_weight_grads(enabled: bool) -> bool:
    prev, _defer_state["enabled"] = _defer_state["enabled"], bool(enabled)
    return prev


def _defer_stream(device, which=0):
    s = _defer_state["streams"].get((device, which))
    if s is None:
        s = _defer_state["streams"][(device, which)] = torch.cuda.Stream(device=device)
    return s


_LAZY_SMALL = os.environ.get("MGNNS_LAZY_SMALL", "0") == "1"
_LSTM_BWD_AFTER_IMAGE = os.environ.get("MGNNS_LSTM_BWD_AFTER_IMAGE", "0") == "1"   # measured: 7.32 vs 7.31 ms (no gain)


def _run_deferred(fn, inputs, params, heavy=False, lazy=False):
    """Enqueue fn() -> list of gradient tensors (one per entry of `params`, None allowed) on the side stream, after
    everything enqueued so far on the current stream; the (param, grad) pairs are applied by join_deferred().
    heavy=True: fn takes the CTA cap of its persistent kernel; only queued here (with an "inputs ready" event) and
    issued by join_deferred() as _HEAVY_PLAN says."""
    dev = inputs[0].device
    cur = torch.cuda.current_stream(dev)
    if heavy:
        _defer_state["heavy"].append((fn, inputs, params, cur.record_event()))
        return
    if lazy and _LAZY_SMALL:
        # issued by join_deferred() behind the first gate: off the SMs while the attention stacks' backward chains run
        _defer_state["lazy"].append((fn, inputs, params, cur.record_event(), _defer_stream(dev, ('side', cur.cuda_stream))))
        return
    # one side stream per origin stream: the four attention stacks' weight gradients do not queue behind each other
    side = _defer_stream(dev, ('side', cur.cuda_stream))
    _defer_state["gate_streams"].append((dev, side))              # joined by join_deferred()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        grads = fn()
    for t in inputs:
        t.record_stream(side)
    for p, g in zip(params, grads):
        if p is not None and g is not None:
            _defer_state["pending"].append((p, g))


def delayed_event(device, key):
    """Event that fires _GATE_DELAY_NS after everything enqueued so far on the current stream has finished (the delay
    runs on a helper stream named by `key`): for releasing a persistent kernel of another stream just BEHIND a
    latency-critical launch that is about to be enqueued here (a few small nodes — weight transposes, counter memsets —
    still sit in front of that launch).  The helper stream is returned too: the caller joins it."""
    ev = torch.cuda.current_stream(device).record_event()
    if _GATE_DELAY_NS <= 0:
        return ev, None
    helper = _defer_stream(device, key)
    helper.wait_event(ev)
    _check(_lib.mgnns_delay_ns(_GATE_DELAY_NS, helper.cuda_stream), "delay")
    return helper.record_event(), helper


def _drop_gate(device):
    """Called right before a latency-critical kernel is enqueued on the current stream (deferral on): the returned
    event fires once everything before that kernel has finished, plus _GATE_DELAY_NS on a helper stream — i.e. just
    after the critical kernel's CTAs have been dispatched."""
    if not (_defer_state["enabled"] and _HEAVY_PLAN):
        return
    ev = torch.cuda.current_stream(device).record_event()
    if _GATE_DELAY_NS > 0:
        helper = _defer_stream(device, 'gate%d' % len(_defer_state["gates"]))
        helper.wait_event(ev)
        _check(_lib.mgnns_delay_ns(_GATE_DELAY_NS, helper.cuda_stream), "delay")
        ev = helper.record_event()
        _defer_state["gate_streams"].append((device, helper))
    _defer_state["gates"].append(ev)


def _before_heavy(device):
    """Work enqueued so far on the current stream must be finished before any heavy deferred job starts."""
    if _defer_state["enabled"] and _HEAVY_PLAN and _HEAVY_AFTER_TEXT:
        _defer_state["before_heavy"].append(torch.cuda.current_stream(device).record_event())


def join_deferred():
    """Issue the gated heavy jobs, join every deferred stream into the current one and accumulate the results into
    .grad (call after loss.backward(), before anything reads the gradients)."""
    heavy, _defer_state["heavy"] = _defer_state["heavy"], []
    gates, _defer_state["gates"] = _defer_state["gates"], []
    before, _defer_state["before_heavy"] = _defer_state["before_heavy"], []
    joined = set(_defer_state["gate_streams"])       # helper streams rejoin even when no heavy job waited on them
    _defer_state["gate_streams"] = []
    for i, (fn, inputs, params, ready) in enumerate(heavy):
        dev = inputs[0].device
        side = _defer_stream(dev, 'heavy%d' % (i % 2))
        side.wait_event(ready)
        for ev in before:
            side.wait_event(ev)
        max_ctas, gate = _HEAVY_PLAN[min(i, len(_HEAVY_PLAN) - 1)]
        if gates and gate >= 0:
            side.wait_event(gates[min(gate, len(gates) - 1)])
        with torch.cuda.stream(side):
            grads = fn(max_ctas)
        for t in inputs:
            t.record_stream(side)
        joined.add((dev, side))
        for p, g in zip(params, grads):
            if p is not None and g is not None:
                _defer_state["pending"].append((p, g))
    lazy, _defer_state["lazy"] = _defer_state["lazy"], []
    for fn, inputs, params, ready, side in lazy:
        dev = inputs[0].device
        side.wait_event(ready)
        if gates:
            side.wait_event(gates[0])
        with torch.cuda.stream(side):
            grads = fn()
        for t in inputs:
            t.record_stream(side)
        joined.add((dev, side))
        for p, g in zip(params, grads):
            if p is not None and g is not None:
                _defer_state["pending"].append((p, g))
    pending, _defer_state["pending"] = _defer_state["pending"], []
    for dev, side in joined:
        torch.cuda.current_stream(dev).wait_stream(side)
    acc_dst, acc_src = [], []
    for p, g in pending:
        g.record_stream(torch.cuda.current_stream(g.device))
        if p.grad is None:
            p.grad = g
        else:
            # in place, as AccumulateGrad does: a never-stepped parameter keeps ONE gradient buffer that accumulates
            # across steps, and a captured graph must keep writing to that same buffer on every replay
            acc_dst.append(p.grad)
            acc_src.append(g)
    if acc_dst:
        with torch.no_grad():
            if len({id(t) for t in acc_dst}) == len(acc_dst):
                torch._foreach_add_(acc_dst, acc_src)       # one multi-tensor launch instead of ~20 tiny ones at the tail
            else:                                           # the same buffer more than once: keep the order explicit
                for d, g in zip(acc_dst, acc_src):
                    d.add_(g)


# ----------------------------------------------------------------------------- helpers
def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


_raw_stream = torch._C._cuda_getCurrentRawStream


def _stream():
    # raw cudaStream_t of torch's current stream on the current device (cheap C call)
    return _raw_stream(torch.cuda.current_device())


def _need_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("mgnns_b200 ops run on CUDA tensors only (got a %s tensor); there is no CPU fallback"
                               % t.device.type)


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise RuntimeError("mgnns_b200: %s must be float32 (got %s)" % (name, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


def _rows2d(t: torch.Tensor, name: str) -> torch.Tensor:
    """2-D fp32 with unit inner stride (arbitrary row stride)."""
    if t.dtype != torch.float32:
        raise RuntimeError("mgnns_b200: %s must be float32 (got %s)" % (name, t.dtype))
    if t.dim() != 2:
        raise RuntimeError("mgnns_b200: %s must be 2-D (got %d-D)" % (name, t.dim()))
    if t.stride(1) != 1 and t.size(1) > 1:
        t = t.contiguous()
    if t.size(1) == 1 and t.stride(0) < 1:
        t = t.contiguous()
    return t


_seed_tensor = None     # optional CUDA int64[1] added to every dropout seed inside the kernels (graph replays)


def set_seed_tensor(t):
    """Install (or clear with None) a device-resident seed offset; GraphedTrainStep advances it once per replay."""
    global _seed_tensor
    prev, _seed_tensor = _seed_tensor, t
    return prev


def _seed_ptr():
    return None if _seed_tensor is None else _seed_tensor.data_ptr()


def new_seed() -> int:
    """Dropout seed drawn from torch's CPU generator (deterministic under torch.manual_seed, no GPU sync)."""
    return int(torch.empty((), dtype=torch.int64).random_().item())


def gemm_raw(transA, transB, M, N, K, A, lda, strideA, B, ldb, strideB, C, ldc, strideC,
             batch=1, reduce=1, accumulate=0, bias=None, act=ACT_NONE, slope=0.0):
    """Direct call of mgnns_gemm_f32 on device pointers held by tensors A, B, C (+ element offsets)."""
    a_t, a_off = A if isinstance(A, tuple) else (A, 0)
    b_t, b_off = B if isinstance(B, tuple) else (B, 0)
    c_t, c_off = C if isinstance(C, tuple) else (C, 0)
    _check(_lib.mgnns_gemm_f32(int(transA), int(transB), M, N, K,
                               a_t.data_ptr() + 4 * a_off, lda, strideA,
                               b_t.data_ptr() + 4 * b_off, ldb, strideB,
                               c_t.data_ptr() + 4 * c_off, ldc, strideC,
                               batch, reduce, accumulate, _ptr(bias), act, float(slope), _stream()), "gemm")


# ----------------------------------------------------------------------------- mm
_LIB.define("mm(Tensor a, Tensor b, Tensor? bias, bool trans_a, bool trans_b, int act, float slope) -> Tensor")


def _mm_shapes(a, b, trans_a, trans_b):
    M, K = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
    Kb, N = (b.shape[1], b.shape[0]) if trans_b else (b.shape[0], b.shape[1])
    if K != Kb:
        raise RuntimeError("mgnns::mm: inner dimensions differ (%d vs %d)" % (K, Kb))
    return M, N, K


_gemm = _lib.mgnns_gemm_f32
_linear_tc = _lib.mgnns_linear_tc
_linear_tc_ws = _lib.mgnns_linear_tc_workspace
_wgrad_tc = _lib.mgnns_wgrad_tc
_TC_MIN_ROWS = int(os.environ.get("MGNNS_TC_MIN_ROWS", "2048"))   # below this the 128-row tiles cannot fill 148 SMs
_gemm_ws = _lib.mgnns_gemm_f32_ws
_f32 = torch.float32
_ws_cache = {}


def _ws_size(M, N, K):
    key = (M, N, K)
    v = _ws_cache.get(key)
    if v is None:
        v = _ws_cache[key] = int(_lib.mgnns_gemm_splitk_workspace(M, N, K))
    return v


def _mm_impl(a, b, bias, trans_a, trans_b, act, slope):
    # hot wrapper (about 150 calls per training step): checks are folded into as few Python operations as possible
    if not (a.is_cuda and b.is_cuda and a.dtype is _f32 and b.dtype is _f32 and a.dim() == 2 and b.dim() == 2):
        _need_cuda(a, b, bias)
        a = _rows2d(a, "a")
        b = _rows2d(b, "b")
    if a.stride(1) != 1 or b.stride(1) != 1:
        a = _rows2d(a, "a")
        b = _rows2d(b, "b")
    if trans_a:
        K, M = a.shape
    else:
        M, K = a.shape
    if trans_b:
        N, Kb = b.shape
    else:
        Kb, N = b.shape
    if K != Kb:
        raise RuntimeError("mgnns::mm: inner dimensions differ (%d vs %d)" % (K, Kb))
    bp = None
    if bias is not None:
        if not (bias.is_cuda and bias.dtype is _f32 and bias.is_contiguous() and bias.numel() == N):
            _need_cuda(bias)
            bias = _f32c(bias, "bias")
            if bias.numel() != N:
                raise RuntimeError("mgnns::mm: bias has %d elements, expected %d" % (bias.numel(), N))
        bp = bias.data_ptr()
    c = torch.empty((M, N), device=a.device, dtype=_f32)
    prec = _PRECISIONS[_precision]
    if (prec is not None and not trans_a and M >= _TC_MIN_ROWS and a.stride(0) % 4 == 0 and b.stride(0) % 4 == 0
            and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0):
        # many-row product: tcgen05 tensor cores (TMA-fed, TMEM accumulators), TF32 or 3xTF32 operands
        w_is_kn = 0 if trans_b else 1
        ldw = b.stride(0)
        wsf = int(_linear_tc_ws(N, K, ldw, w_is_kn, prec))
        ws = torch.empty((max(wsf, 4),), device=a.device, dtype=_f32)
        with _timed("linear_tc"):
            rc = _linear_tc(a.data_ptr(), a.stride(0), b.data_ptr(), ldw, w_is_kn, bp, act, slope, M, N, K, prec,
                            ws.data_ptr(), wsf, c.data_ptr(), N, _raw_stream(a.device.index))
        if rc:
            _check(rc, "linear_tc")
        return c
    if (prec is not None and trans_a and not trans_b and K >= _TC_MIN_ROWS and a.stride(0) % 4 == 0 and b.stride(0) % 4 == 0
            and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0 and bp is None and act == ACT_NONE):
        # weight-gradient shape (reduction over many rows): tcgen05 kernel with split-K atomics
        with _timed("wgrad_tc"):
            rc = _wgrad_tc(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), M, N, K, prec, c.data_ptr(), N,
                           _raw_stream(a.device.index))
        if rc:
            _check(rc, "wgrad_tc")
        return c
    ws_floats = 0 if trans_a else _ws_size(M, N, K)
    if ws_floats:
        # small forward-shaped product: deterministic two-pass split-K through a scratch buffer
        ws = torch.empty((ws_floats,), device=a.device, dtype=_f32)
        rc = _gemm_ws(trans_a, trans_b, M, N, K, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0),
                      c.data_ptr(), N, bp, act, slope, ws.data_ptr(), ws_floats, _raw_stream(a.device.index))
    else:
        rc = _gemm(trans_a, trans_b, M, N, K, a.data_ptr(), a.stride(0), 0, b.data_ptr(), b.stride(0), 0,
                   c.data_ptr(), N, 0, 1, 1, 0, bp, act, slope, _raw_stream(a.device.index))
    if rc:
        _check(rc, "gemm")
    return c


def _mm_fake(a, b, bias, trans_a, trans_b, act, slope):
    M, N, _ = _mm_shapes(a, b, trans_a, trans_b)
    return a.new_empty((M, N))


_LIB.impl("mm", _mm_impl, "CUDA")
torch.library.register_fake("mgnns::mm", _mm_fake)


def act_bwd(y, g, act, slope):
    if act == ACT_NONE:
        return g
    y = _f32c(y, "y")
    g = _f32c(g, "g")
    out = torch.empty_like(g)
    _check(_lib.mgnns_act_bwd_f32(y.data_ptr(), g.data_ptr(), out.data_ptr(), g.numel(), act, float(slope), _stream()),
           "act_bwd")
    return out


def colsum(x2d):
    x2d = _rows2d(x2d, "x")
    out = torch.zeros((x2d.shape[1],), device=x2d.device, dtype=torch.float32)
    _check(_lib.mgnns_colsum_f32(x2d.data_ptr(), x2d.shape[0], x2d.shape[1], x2d.stride(0), out.data_ptr(), _stream()),
           "colsum")
    return out


def _mm_setup(ctx, inputs, output):
    a, b, bias, trans_a, trans_b, act, slope = inputs
    ctx.save_for_backward(a, b, output if act != ACT_NONE else None)
    ctx.cfg = (trans_a, trans_b, act, slope, bias is not None)
    ctx.leaves = (b, bias)                 # the parameter objects themselves (deferred weight gradients)


def _mm_backward(ctx, g):
    a, b, y = ctx.saved_tensors
    ta, tb, act, slope, has_bias = ctx.cfg
    g = act_bwd(y, g, act, slope) if act != ACT_NONE else _f32c(g, "grad")
    ga = gb = gbias = None
    mm = _mm_impl            # backward needs no autograd graph: skip the dispatcher
    if ctx.needs_input_grad[0]:
        ga = mm(g, b, None, False, not tb, ACT_NONE, 0.0) if not ta else mm(b, g, None, tb, True, ACT_NONE, 0.0)
    need_w, need_bias = ctx.needs_input_grad[1], has_bias and ctx.needs_input_grad[2]

    def weight_grads():
        gw_ = gbias_ = None
        if need_w:
            gw_ = mm(a, g, None, not ta, False, ACT_NONE, 0.0) if not tb else mm(g, a, None, True, ta, ACT_NONE, 0.0)
        if need_bias:
            gbias_ = colsum(g)
        return [gw_, gbias_]

    w_leaf, bias_leaf = ctx.leaves
    if ((need_w or need_bias) and _defer_state["enabled"] and _DEFER_SMALL and w_leaf.is_leaf
            and (bias_leaf is None or bias_leaf.is_leaf)):
        # a parameter's gradient is not needed before the optimizer: off the backward chain, onto the side stream
        _run_deferred(weight_grads, (g, a), (w_leaf if need_w else None, bias_leaf if need_bias else None), lazy=True)
    elif need_w or need_bias:
        gb, gbias = weight_grads()
    return ga, gb, gbias, None, None, None, None


torch.library.register_autograd("mgnns::mm", _mm_backward, setup_context=_mm_setup)


def linear(x, weight, bias=None, act=ACT_NONE, slope=0.0):
    """act(x @ weight.T + bias) with nn.Linear's [out,in] weight; x may have leading dims."""
    lead = x.shape[:-1]
    y = torch.ops.mgnns.mm(x.reshape(-1, x.shape[-1]), weight, bias, False, True, act, slope)
    return y.reshape(*lead, weight.shape[0])


def matmul_nn(x, weight, bias=None, act=ACT_NONE, slope=0.0):
    """act(x @ weight + bias) with a [in,out] weight (GraphConvolution layout)."""
    lead = x.shape[:-1]
    y = torch.ops.mgnns.mm(x.reshape(-1, x.shape[-1]), weight, bias, False, False, act, slope)
    return y.reshape(*lead, weight.shape[1])


# ----------------------------------------------------------------------------- head_mm
# mode 0: out[b, h*D:(h+1)*D]   = x[b, h*dk:(h+1)*dk] @ w[h*dk:(h+1)*dk, :]        (x [B,H*dk], w [H*dk, D])
# mode 1: out[b, h*dv:(h+1)*dv] = x[b, h*D:(h+1)*D]   @ w[h*dv:(h+1)*dv, :].T      (x [B,H*D],  w [H*dv, D])
_LIB.define("head_mm(Tensor x, Tensor w, int heads, int mode) -> Tensor")


def _head_mm_impl(x, w, heads, mode):
    _need_cuda(x, w)
    x = _f32c(x, "x")
    w = _f32c(w, "w")
    B = x.shape[0]
    D = w.shape[1]
    dk = w.shape[0] // heads
    if w.shape[0] != heads * dk:
        raise RuntimeError("mgnns::head_mm: weight rows not divisible by heads")
    if mode == 0:
        if x.shape[1] != heads * dk:
            raise RuntimeError("mgnns::head_mm: x has %d columns, expected %d" % (x.shape[1], heads * dk))
        out = torch.empty((B, heads * D), device=x.device, dtype=torch.float32)
        gemm_raw(0, 0, B, D, dk, x, heads * dk, dk, w, D, dk * D, out, heads * D, D, batch=heads)
    else:
        if x.shape[1] != heads * D:
            raise RuntimeError("mgnns::head_mm: x has %d columns, expected %d" % (x.shape[1], heads * D))
        out = torch.empty((B, heads * dk), device=x.device, dtype=torch.float32)
        gemm_raw(0, 1, B, dk, D, x, heads * D, D, w, D, dk * D, out, heads * dk, dk, batch=heads)
    return out


def _head_mm_fake(x, w, heads, mode):
    D = w.shape[1]
    dk = w.shape[0] // heads
    return x.new_empty((x.shape[0], heads * (D if mode == 0 else dk)))


_LIB.impl("head_mm", _head_mm_impl, "CUDA")
torch.library.register_fake("mgnns::head_mm", _head_mm_fake)


def _head_mm_setup(ctx, inputs, output):
    x, w, heads, mode = inputs
    ctx.save_for_backward(x, w)
    ctx.cfg = (heads, mode)
    ctx.leaves = (w,)


def _head_mm_backward(ctx, g):
    x, w = ctx.saved_tensors
    heads, mode = ctx.cfg
    g = _f32c(g, "grad")
    x = _f32c(x, "x")
    B = x.shape[0]
    D = w.shape[1]
    dk = w.shape[0] // heads
    gx = gw = None
    if ctx.needs_input_grad[0]:
        gx = torch.ops.mgnns.head_mm(g, w, heads, 1 - mode)
    def weight_grads():
        gw_ = torch.empty_like(w)
        if mode == 0:   # gw_h [dk, D] = x_h^T @ g_h
            gemm_raw(1, 0, dk, D, B, x, heads * dk, dk, g, heads * D, D, gw_, D, dk * D, batch=heads)
        else:           # gw_h [dv, D] = g_h^T @ x_h
            gemm_raw(1, 0, dk, D, B, g, heads * dk, dk, x, heads * D, D, gw_, D, dk * D, batch=heads)
        return [gw_]

    if ctx.needs_input_grad[1]:
        if _defer_state["enabled"] and _DEFER_SMALL and ctx.leaves[0].is_leaf:
            _run_deferred(weight_grads, (g, x), ctx.leaves, lazy=True)
        else:
            gw = weight_grads()[0]
    return gx, gw, None, None


torch.library.register_autograd("mgnns::head_mm", _head_mm_backward, setup_context=_head_mm_setup)


# ----------------------------------------------------------------------------- spmm
_LIB.define("spmm_csr(Tensor rowptr, Tensor col, Tensor val, Tensor x, Tensor t_rowptr, Tensor t_col, "
            "Tensor t_val, int n_rows) -> Tensor")


_HUB_MIN_BATCH = int(os.environ.get("MGNNS_SPMM_HUB_MIN_BATCH", "16"))
_HUB_MIN_NNZ = int(os.environ.get("MGNNS_SPMM_HUB_MIN_NNZ", "100000"))
_hub_plans = {}


def _hub_plan(rowptr, col, val, n_rows, n_cols, F, ldx):
    """Execution plan of a CSR matrix for mgnns_spmm_hub_f32, built on the host once per (matrix, feature width) —
    one device->host copy of the CSR arrays — and cached on the identity and version of the three tensors."""
    key = (rowptr.data_ptr(), col.data_ptr(), val.data_ptr(), rowptr._version, col._version, val._version, n_cols, F, ldx)
    plan = _hub_plans.get(key)
    if plan is None:
        from .api.graph_util import HubSpmmPlan, hub_plan_arrays
        sms = torch.cuda.get_device_properties(x_dev := rowptr.device).multi_processor_count
        arrays = hub_plan_arrays(rowptr.cpu().numpy(), col.cpu().numpy(), val.cpu().numpy(), n_cols, ldx, F,
                                 int(_lib.mgnns_spmm_hub_capacity(F)), sms, seg_edges=64)
        plan = HubSpmmPlan(arrays, ldx, F, x_dev)
        if len(_hub_plans) > 16:
            _hub_plans.clear()
        _hub_plans[key] = plan
    return plan


def _spmm_raw(rowptr, col, val, x, n_rows):
    _need_cuda(rowptr, col, val, x)
    x = _f32c(x, "x")
    if x.dim() == 2:
        batch, n_cols, F = 1, x.shape[0], x.shape[1]
        y = torch.empty((n_rows, F), device=x.device, dtype=torch.float32)
    elif x.dim() == 3:
        batch, n_cols, F = x.shape
        y = torch.empty((batch, n_rows, F), device=x.device, dtype=torch.float32)
    else:
        raise RuntimeError("mgnns::spmm_csr: x must be [N,F] or [B,N,F]")
    if rowptr.dtype != torch.int32 or col.dtype != torch.int32 or val.dtype != torch.float32:
        raise RuntimeError("mgnns::spmm_csr: CSR arrays must be int32/int32/float32")
    if rowptr.numel() != n_rows + 1:
        raise RuntimeError("mgnns::spmm_csr: rowptr has %d entries, expected %d" % (rowptr.numel(), n_rows + 1))
    if (batch >= _HUB_MIN_BATCH and col.numel() >= _HUB_MIN_NNZ and F % 4 == 0 and F <= 512
            and x.data_ptr() % 16 == 0 and os.environ.get("MGNNS_SPMM_HUB", "0") == "1"):
        # Opt-in: persistent kernel with the hub neighbour rows staged in shared memory.  Measured on cfg 2 (B200):
        # 11.97 ms against 11.33 ms for the plain kernel — the gather is bound by the L1 global-load return path
        # (64 B/clk/SM: 17.6 of 18.2 TB/s), which the plain kernel already saturates; moving the 31 % hub edges to
        # shared memory (128 B/clk) is offset by this kernel's lower occupancy.
        # MGNNS_SPMM_PAD=1 copies X into rows padded to 128 bytes first (rows of 300 floats start at 16-byte aligned
        # addresses): measured, it buys nothing — 13.4 ms with the copy — so alignment is not what limits the gather.
        ldx = F
        if os.environ.get("MGNNS_SPMM_PAD", "0") == "1" and F % 32 != 0:
            ldx = (F + 31) // 32 * 32
            with _timed("spmm_pad"):
                xp = torch.empty((batch, n_cols, ldx), device=x.device, dtype=torch.float32)
                xp[:, :, :F].copy_(x)
            x = xp
        plan = _hub_plan(rowptr, col, val, n_rows, n_cols, F, ldx)
        with _timed("spmm_hub"):
            _check(_lib.mgnns_spmm_hub_f32(x.data_ptr(), ldx, n_cols * ldx, y.data_ptr(), F, n_rows * F, F, batch,
                                           plan.hub_cols.data_ptr(), plan.n_hub, plan.chunk_seg_ptr.data_ptr(), plan.n_chunks,
                                           plan.segs.data_ptr(), plan.edges.data_ptr(), plan.multi_rows.data_ptr(),
                                           plan.n_multi, _stream()), "spmm_hub")
        return y
    for b0 in range(0, batch, 65535):
        nb = min(65535, batch - b0)
        with _timed("spmm_csr"):
            _check(_lib.mgnns_spmm_csr_f32(n_rows, rowptr.data_ptr(), col.data_ptr(), val.data_ptr(),
                                           x.data_ptr() + 4 * b0 * n_cols * F, F, n_cols * F,
                                           y.data_ptr() + 4 * b0 * n_rows * F, F, n_rows * F,
                                           F, nb, _stream()), "spmm_csr")
    return y


def _spmm_impl(rowptr, col, val, x, t_rowptr, t_col, t_val, n_rows):
    return _spmm_raw(rowptr, col, val, x, n_rows)


def _spmm_fake(rowptr, col, val, x, t_rowptr, t_col, t_val, n_rows):
    shape = list(x.shape)
    shape[-2] = n_rows
    return x.new_empty(shape)


_LIB.impl("spmm_csr", _spmm_impl, "CUDA")
torch.library.register_fake("mgnns::spmm_csr", _spmm_fake)


def _spmm_setup(ctx, inputs, output):
    rowptr, col, val, x, t_rowptr, t_col, t_val, n_rows = inputs
    ctx.save_for_backward(t_rowptr, t_col, t_val)
    ctx.n_cols = x.shape[-2]


def _spmm_backward(ctx, g):
    t_rowptr, t_col, t_val = ctx.saved_tensors
    gx = None
    if ctx.needs_input_grad[3]:
        gx = _spmm_raw(t_rowptr, t_col, t_val, g, ctx.n_cols)
    return None, None, None, gx, None, None, None, None


torch.library.register_autograd("mgnns::spmm_csr", _spmm_backward, setup_context=_spmm_setup)


# ----------------------------------------------------------------------------- fused graph-convolution layer
_LIB.define("gcn_fused(Tensor x, Tensor weight, Tensor? bias, Tensor tile_seg_ptr, Tensor segs, Tensor edges, "
            "Tensor tile_rows, Tensor tile_multi_ptr, Tensor multi_rows, int n_tiles, int n_rows, int act, "
            "float slope) -> Tensor")
_gcn_fused = _lib.mgnns_gcn_fused_tc
_gcn_fused_ws = _lib.mgnns_gcn_fused_workspace


def gcn_fused_ok(x, weight) -> bool:
    """Shapes / mode the fused layer kernel covers (forward only; callers that need gradients use spmm + mm).
    Opt-in (MGNNS_GCN_FUSED=1): measured on cfg 2 the fused kernel moves 1.0x the algorithmic HBM bytes (the
    two-kernel path 2.1x) but is slower, because one 768-thread CTA per SM cannot keep as many neighbour-row loads
    in flight as the stand-alone SpMM."""
    K, N = weight.shape
    return (_PRECISIONS[_precision] is not None and x.dim() == 3 and x.is_cuda and x.dtype is _f32
            and K % 4 == 0 and N % 32 == 0 and N <= 512 and K <= N and x.shape[2] == K
            and os.environ.get("MGNNS_GCN_FUSED", "0") == "1")


def _gcn_fused_impl(x, weight, bias, tile_seg_ptr, segs, edges, tile_rows, tile_multi_ptr, multi_rows, n_tiles, n_rows,
                    act, slope):
    _need_cuda(x, weight, bias, tile_seg_ptr, segs, edges, tile_rows, tile_multi_ptr, multi_rows)
    prec = _PRECISIONS[_precision]
    if prec is None:
        raise RuntimeError("mgnns::gcn_fused runs on the tensor cores (precision tf32x3 or tf32), not in fp32 mode")
    if x.dim() != 3:
        raise RuntimeError("mgnns::gcn_fused: x must be [batch, nodes, in_features]")
    x = _f32c(x, "x")
    weight = _f32c(weight, "weight")
    batch, _, K = x.shape
    if weight.shape[0] != K:
        raise RuntimeError("mgnns::gcn_fused: weight has %d rows, x has %d features" % (weight.shape[0], K))
    N = weight.shape[1]
    for t in (tile_seg_ptr, segs, edges, tile_rows, tile_multi_ptr, multi_rows):
        if t.dtype != torch.int32 or not t.is_contiguous():
            raise RuntimeError("mgnns::gcn_fused: plan arrays must be contiguous int32")
    bp = None
    if bias is not None:
        bias = _f32c(bias.reshape(-1), "bias")
        if bias.numel() != N:
            raise RuntimeError("mgnns::gcn_fused: bias has %d elements, expected %d" % (bias.numel(), N))
        bp = bias.data_ptr()
    y = torch.empty((batch, n_rows, N), device=x.device, dtype=_f32)
    wsf = int(_gcn_fused_ws(N, K, prec))
    ws = torch.empty((max(wsf, 4),), device=x.device, dtype=_f32)
    with _timed("gcn_fused"):
        rc = _gcn_fused(x.data_ptr(), x.stride(0), batch, tile_seg_ptr.data_ptr(), segs.data_ptr(), edges.data_ptr(),
                        tile_rows.data_ptr(), tile_multi_ptr.data_ptr(), multi_rows.data_ptr(), n_tiles,
                        weight.data_ptr(), weight.stride(0), bp, act, slope, K, N, prec, ws.data_ptr(), wsf,
                        y.data_ptr(), N, n_rows * N, _raw_stream(x.device.index))
    if rc:
        _check(rc, "gcn_fused")
    return y


def _gcn_fused_fake(x, weight, bias, tile_seg_ptr, segs, edges, tile_rows, tile_multi_ptr, multi_rows, n_tiles, n_rows,
                    act, slope):
    return x.new_empty((x.shape[0], n_rows, weight.shape[1]))


_LIB.impl("gcn_fused", _gcn_fused_impl, "CUDA")
torch.library.register_fake("mgnns::gcn_fused", _gcn_fused_fake)


def gcn_fused(plan, x, weight, bias=None, act=ACT_NONE, slope=0.0):
    """act((Â·x)·weight + bias) in one kernel; `plan` = CSRAdjacency.fused_plan(x.shape[2])."""
    return torch.ops.mgnns.gcn_fused(x, weight, bias, plan.tile_seg_ptr, plan.segs, plan.edges, plan.tile_rows,
                                     plan.tile_multi_ptr, plan.multi_rows, plan.n_tiles, plan.n_rows, act, float(slope))


# ----------------------------------------------------------------------------- dense -> CSR
def dense_to_csr(adj: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """[n,m] fp32 CUDA matrix -> (rowptr int32[n+1], col int32[nnz], val fp32[nnz]); one host sync for nnz."""
    _need_cuda(adj)
    adj = _rows2d(adj, "adj")
    n, m = adj.shape
    nnz_row = torch.empty((n,), device=adj.device, dtype=torch.int32)
    rowptr = torch.empty((n + 1,), device=adj.device, dtype=torch.int32)
    s = _stream()
    _check(_lib.mgnns_dense_row_nnz_f32(adj.data_ptr(), n, m, adj.stride(0), nnz_row.data_ptr(), s), "dense_row_nnz")
    _check(_lib.mgnns_exclusive_scan_i32(nnz_row.data_ptr(), rowptr.data_ptr(), n, s), "scan")
    nnz = int(rowptr[-1].item())
    col = torch.empty((max(nnz, 1),), device=adj.device, dtype=torch.int32)[:nnz]
    val = torch.empty((max(nnz, 1),), device=adj.device, dtype=torch.float32)[:nnz]
    _check(_lib.mgnns_dense_fill_csr_f32(adj.data_ptr(), n, m, adj.stride(0), rowptr.data_ptr(),
                                         col.data_ptr(), val.data_ptr(), s), "dense_fill_csr")
    return rowptr, col, val


# ----------------------------------------------------------------------------- text max-aggregation
_LIB.define("text_maxagg(Tensor doc_ids, Tensor node_hidden, Tensor edge_w, Tensor pmi_rowptr, Tensor pmi_col, "
            "Tensor? pmi_eid, int ngram, int max_length, bool apply_relu) -> Tensor")


def _text_args(doc_ids, node_hidden, edge_w, pmi_rowptr, pmi_col, pmi_eid):
    _need_cuda(doc_ids, node_hidden, edge_w, pmi_rowptr, pmi_col, pmi_eid)
    if doc_ids.dtype != torch.int64 or doc_ids.dim() != 2:
        raise RuntimeError("mgnns::text_maxagg: doc_ids must be int64 [B,L]")
    doc_ids = doc_ids.contiguous()
    node_hidden = _f32c(node_hidden, "node_hidden")
    edge_w = _f32c(edge_w, "edge_w")
    if pmi_rowptr.dtype != torch.int32 or pmi_col.dtype != torch.int32:
        raise RuntimeError("mgnns::text_maxagg: PMI CSR arrays must be int32")
    if pmi_rowptr.numel() != node_hidden.shape[0] + 1:
        raise RuntimeError("mgnns::text_maxagg: pmi_rowptr must have V+1 entries")
    if pmi_eid is not None and pmi_eid.dtype != torch.int32:
        raise RuntimeError("mgnns::text_maxagg: pmi_eid must be int32")
    return doc_ids, node_hidden, edge_w


def _text_impl(doc_ids, node_hidden, edge_w, pmi_rowptr, pmi_col, pmi_eid, ngram, max_length, apply_relu):
    doc_ids, node_hidden, edge_w = _text_args(doc_ids, node_hidden, edge_w, pmi_rowptr, pmi_col, pmi_eid)
    B, L = doc_ids.shape
    V, F = node_hidden.shape
    out = torch.empty((B, F), device=node_hidden.device, dtype=torch.float32)
    _check(_lib.mgnns_text_maxagg_fwd(doc_ids.data_ptr(), B, L, max_length, ngram, node_hidden.data_ptr(), V, F,
                                      edge_w.data_ptr(), edge_w.numel(), pmi_rowptr.data_ptr(), pmi_col.data_ptr(),
                                      _ptr(pmi_eid), int(apply_relu), out.data_ptr(), _stream()), "text_maxagg_fwd")
    return out


def _text_fake(doc_ids, node_hidden, edge_w, pmi_rowptr, pmi_col, pmi_eid, ngram, max_length, apply_relu):
    return node_hidden.new_empty((doc_ids.shape[0], node_hidden.shape[1]))


_LIB.impl("text_maxagg", _text_impl, "CUDA")
torch.library.register_fake("mgnns::text_maxagg", _text_fake)


def _text_setup(ctx, inputs, output):
    doc_ids, node_hidden, edge_w, pmi_rowptr, pmi_col, pmi_eid, ngram, max_length, apply_relu = inputs
    ctx.save_for_backward(doc_ids, node_hidden, edge_w, pmi_rowptr, pmi_col, pmi_eid, output)
    ctx.cfg = (ngram, max_length, apply_relu)


def _text_backward(ctx, g):
    doc_ids, node_hidden, edge_w, pmi_rowptr, pmi_col, pmi_eid, out = ctx.saved_tensors
    ngram, max_length, apply_relu = ctx.cfg
    doc_ids, node_hidden, edge_w_c = _text_args(doc_ids, node_hidden, edge_w, pmi_rowptr, pmi_col, pmi_eid)
    g = _f32c(g, "grad")
    B, L = doc_ids.shape
    V, F = node_hidden.shape
    g_h = torch.zeros_like(node_hidden)
    g_w = torch.zeros((edge_w_c.numel(),), device=g.device, dtype=torch.float32)
    _check(_lib.mgnns_text_maxagg_bwd(doc_ids.data_ptr(), B, L, max_length, ngram, node_hidden.data_ptr(), V, F,
                                      edge_w_c.data_ptr(), edge_w_c.numel(), pmi_rowptr.data_ptr(),
                                      pmi_col.data_ptr(), _ptr(pmi_eid), int(apply_relu), out.data_ptr(),
                                      g.data_ptr(), g_h.data_ptr(), g_w.data_ptr(), _stream()), "text_maxagg_bwd")
    # The text-GCN backward becomes ready together with the image-bank weight gradients (both hang off the
    # image-query stacks' backward) and needs few SMs for ~0.13 ms; started behind two persistent kernels that hold
    # every SM it ran last, 1.6 ms later, as the tail of the whole step.  Heavy deferred jobs therefore go behind it.
    _before_heavy(g.device)
    return None, g_h, g_w.view_as(edge_w), None, None, None, None, None, None


torch.library.register_autograd("mgnns::text_maxagg", _text_backward, setup_context=_text_setup)


# ----------------------------------------------------------------------------- single-query attention core
_LIB.define("attn_q1(Tensor u, Tensor bank, Tensor? mask, float scale, float p_drop, int seed) "
            "-> (Tensor, Tensor, Tensor, Tensor)")


def _attn_check(u, bank, mask):
    _need_cuda(u, bank, mask)
    u = _f32c(u, "u")
    bank = _f32c(bank, "bank")
    if u.dim() != 3 or bank.dim() != 3 or u.shape[0] != bank.shape[0] or u.shape[2] != bank.shape[2]:
        raise RuntimeError("mgnns::attn_q1: expected u [B,H,D] and bank [B,L,D]")
    if mask is not None:
        mask = _f32c(mask, "mask")
        if mask.shape != bank.shape[:2]:
            raise RuntimeError("mgnns::attn_q1: mask must be [B,L]")
    return u, bank, mask


_attn_tc_cache = {}


def attn_uses_tensor_cores(H, L, D) -> bool:
    """True when the mma.sync (3xTF32) attention kernels cover the shape and MGNNS_ATTN is not 'scalar'."""
    if os.environ.get("MGNNS_ATTN", "tc") == "scalar":
        return False
    key = (H, L, D)
    v = _attn_tc_cache.get(key)
    if v is None:
        v = _attn_tc_cache[key] = bool(_lib.mgnns_attn_q1_tc_supported(H, L, D))
    return v


def _attn_impl(u, bank, mask, scale, p_drop, seed):
    u, bank, mask = _attn_check(u, bank, mask)
    B, H, D = u.shape
    L = bank.shape[1]
    fwd = _lib.mgnns_attn_q1_tc_fwd if attn_uses_tensor_cores(H, L, D) else _lib.mgnns_attn_q1_fwd
    ctx = torch.empty((B, H, D), device=u.device, dtype=torch.float32)
    attn = torch.empty((H * B, 1, L), device=u.device, dtype=torch.float32)
    psum = torch.empty((B, H), device=u.device, dtype=torch.float32)
    lse = torch.empty((B, H), device=u.device, dtype=torch.float32)
    with _timed("attn_q1_fwd"):
        _check(fwd(u.data_ptr(), bank.data_ptr(), _ptr(mask), B, H, L, D, float(scale),
                   float(p_drop), seed & 0xFFFFFFFFFFFFFFFF, _seed_ptr(), ctx.data_ptr(), attn.data_ptr(),
                   psum.data_ptr(), lse.data_ptr(), _stream()), "attn_q1_fwd")
    return ctx, attn, psum, lse


def _attn_fake(u, bank, mask, scale, p_drop, seed):
    B, H, D = u.shape
    L = bank.shape[1]
    return u.new_empty((B, H, D)), u.new_empty((H * B, 1, L)), u.new_empty((B, H)), u.new_empty((B, H))


_LIB.impl("attn_q1", _attn_impl, "CUDA")
torch.library.register_fake("mgnns::attn_q1", _attn_fake)


def _attn_setup(ctx, inputs, output):
    u, bank, mask, scale, p_drop, seed = inputs
    ctx.save_for_backward(u, bank, mask, output[3])
    ctx.cfg = (scale, p_drop, seed)
    ctx.set_materialize_grads(False)          # unused outputs arrive as None in backward (see the check there)


def _attn_backward(ctx, g_ctx, g_attn, g_psum, g_lse):
    if g_attn is not None or g_lse is not None:
        # the attention weights / log-sum-exp are returned for inspection (the reference returns attn too, submodules.py:90)
        # but no gradient formula flows through them here; a loss term built on them must fail loudly, not get zeros
        raise NotImplementedError("mgnns::attn_q1: gradients through the returned attention weights / lse are not implemented")
    u, bank, mask, lse = ctx.saved_tensors
    scale, p_drop, seed = ctx.cfg
    u, bank, mask = _attn_check(u, bank, mask)
    B, H, D = u.shape
    L = bank.shape[1]
    g_ctx = torch.zeros_like(u) if g_ctx is None else _f32c(g_ctx, "grad_ctx")
    g_psum = None if g_psum is None else _f32c(g_psum, "grad_psum")
    gu = torch.empty_like(u)
    gbank = torch.empty_like(bank)
    bwd = _lib.mgnns_attn_q1_tc_bwd if attn_uses_tensor_cores(H, L, D) else _lib.mgnns_attn_q1_bwd
    with _timed("attn_q1_bwd"):
        _check(bwd(u.data_ptr(), bank.data_ptr(), _ptr(mask), lse.data_ptr(), g_ctx.data_ptr(),
                   _ptr(g_psum), B, H, L, D, float(scale), float(p_drop),
                   seed & 0xFFFFFFFFFFFFFFFF, _seed_ptr(), gu.data_ptr(), gbank.data_ptr(), _stream()),
               "attn_q1_bwd")
    return gu, gbank, None, None, None, None


torch.library.register_autograd("mgnns::attn_q1", _attn_backward, setup_context=_attn_setup)


# ----------------------------------------------------------------------------- label attention
_LIB.define("label_attn(Tensor q, Tensor kv, int heads, float inv_scale, float p_drop, int seed) -> Tensor")


def _label_check(q, kv, heads):
    _need_cuda(q, kv)
    q = _f32c(q, "q")
    kv = _f32c(kv, "kv")
    HD = q.shape[1]
    if kv.dim() != 2 or kv.shape[1] != 2 * HD or HD % heads != 0:
        raise RuntimeError("mgnns::label_attn: expected q [C,HD], kv [B,2*HD], HD divisible by heads")
    return q, kv, HD


def _label_impl(q, kv, heads, inv_scale, p_drop, seed):
    q, kv, HD = _label_check(q, kv, heads)
    B, C = kv.shape[0], q.shape[0]
    out = torch.empty((B, C, HD), device=q.device, dtype=torch.float32)
    _check(_lib.mgnns_label_attn_fwd(q.data_ptr(), kv.data_ptr(), kv.data_ptr() + 4 * HD, 2 * HD, B, C, heads,
                                     HD // heads, float(inv_scale), float(p_drop), seed & 0xFFFFFFFFFFFFFFFF, _seed_ptr(),
                                     out.data_ptr(), _stream()), "label_attn_fwd")
    return out


def _label_fake(q, kv, heads, inv_scale, p_drop, seed):
    return q.new_empty((kv.shape[0], q.shape[0], q.shape[1]))


_LIB.impl("label_attn", _label_impl, "CUDA")
torch.library.register_fake("mgnns::label_attn", _label_fake)


def _label_setup(ctx, inputs, output):
    q, kv, heads, inv_scale, p_drop, seed = inputs
    ctx.save_for_backward(q, kv)
    ctx.cfg = (heads, inv_scale, p_drop, seed)


def _label_backward(ctx, g):
    q, kv = ctx.saved_tensors
    heads, inv_scale, p_drop, seed = ctx.cfg
    q, kv, HD = _label_check(q, kv, heads)
    g = _f32c(g, "grad")
    B, C = kv.shape[0], q.shape[0]
    gq = torch.zeros_like(q)
    gkv = torch.empty_like(kv)
    _check(_lib.mgnns_label_attn_bwd(q.data_ptr(), kv.data_ptr(), kv.data_ptr() + 4 * HD, 2 * HD, B, C, heads,
                                     HD // heads, float(inv_scale), float(p_drop), seed & 0xFFFFFFFFFFFFFFFF, _seed_ptr(),
                                     g.data_ptr(), gq.data_ptr(), gkv.data_ptr(), gkv.data_ptr() + 4 * HD, 2 * HD,
                                     _stream()), "label_attn_bwd")
    return gq, gkv, None, None, None, None


torch.library.register_autograd("mgnns::label_attn", _label_backward, setup_context=_label_setup)


# ----------------------------------------------------------------------------- add + LayerNorm
_LIB.define("add_layernorm(Tensor x, Tensor? res, Tensor gamma, Tensor beta, float eps) -> Tensor")


def _ln_impl(x, res, gamma, beta, eps):
    _need_cuda(x, res, gamma, beta)
    x = _f32c(x, "x")
    res = None if res is None else _f32c(res, "res")
    gamma = _f32c(gamma, "gamma")
    beta = _f32c(beta, "beta")
    D = x.shape[-1]
    if gamma.numel() != D or beta.numel() != D or (res is not None and res.shape != x.shape):
        raise RuntimeError("mgnns::add_layernorm: shape mismatch")
    y = torch.empty_like(x)
    _check(_lib.mgnns_add_layernorm_fwd(x.data_ptr(), _ptr(res), gamma.data_ptr(), beta.data_ptr(),
                                        x.numel() // D, D, float(eps), y.data_ptr(), _stream()), "add_layernorm_fwd")
    return y


def _ln_fake(x, res, gamma, beta, eps):
    return torch.empty_like(x)


_LIB.impl("add_layernorm", _ln_impl, "CUDA")
torch.library.register_fake("mgnns::add_layernorm", _ln_fake)


def _ln_setup(ctx, inputs, output):
    x, res, gamma, beta, eps = inputs
    ctx.save_for_backward(x, res, gamma)
    ctx.eps = eps


def _ln_backward(ctx, g):
    x, res, gamma = ctx.saved_tensors
    x = _f32c(x, "x")
    res = None if res is None else _f32c(res, "res")
    gamma = _f32c(gamma, "gamma")
    g = _f32c(g, "grad")
    D = x.shape[-1]
    gz = torch.empty_like(x)
    gg = torch.zeros_like(gamma)
    gb = torch.zeros_like(gamma)
    _check(_lib.mgnns_add_layernorm_bwd(x.data_ptr(), _ptr(res), gamma.data_ptr(), g.data_ptr(), x.numel() // D, D,
                                        float(ctx.eps), gz.data_ptr(), gg.data_ptr(), gb.data_ptr(), _stream()),
           "add_layernorm_bwd")
    return gz, (gz if res is not None else None), gg, gb, None


torch.library.register_autograd("mgnns::add_layernorm", _ln_backward, setup_context=_ln_setup)


def rowmax(x2d: torch.Tensor):
    """Row-wise max with first-index arg-max of a contiguous [rows, P] fp32 CUDA matrix (no autograd);
    the kernel behind the 14x14 global max pool (ref: nn.MaxPool2d(14,14), model:302,:454,:486)."""
    _need_cuda(x2d)
    x2d = _f32c(x2d, "x")
    rows, P = x2d.shape
    pooled = torch.empty((rows,), device=x2d.device, dtype=torch.float32)
    argmax = torch.empty((rows,), device=x2d.device, dtype=torch.int32)
    _check(_lib.mgnns_rowmax_f32(x2d.data_ptr(), rows, P, pooled.data_ptr(), argmax.data_ptr(), _stream()), "rowmax")
    return pooled, argmax


# ----------------------------------------------------------------------------- image bank (+ global max pool)
_LIB.define("imgbank(Tensor fmap, Tensor weight, Tensor bias) -> (Tensor, Tensor, Tensor)")


def _imgbank_check(fmap, weight, bias):
    _need_cuda(fmap, weight, bias)
    fmap = _f32c(fmap, "fmap")
    weight = _f32c(weight, "weight")
    bias = _f32c(bias, "bias")
    if fmap.dim() == 4:
        fmap = fmap.reshape(fmap.shape[0], fmap.shape[1], -1)
    if fmap.dim() != 3 or weight.dim() != 2 or weight.shape[1] != fmap.shape[1] or bias.numel() != weight.shape[0]:
        raise RuntimeError("mgnns::imgbank: expected fmap [B,C,H,W], weight [O,C], bias [O]")
    return fmap, weight, bias


# SMs the persistent image-bank kernels leave free while the model runs its chains on several streams (their CTAs
# take a whole SM each, so a grid on every SM stalls every small kernel of the other streams for ~0.8 ms).  Measured
# on the B=512 training step: 0 -> 7.46 ms, 4 / 8 / 16 -> 7.56, 24 -> 7.62: the small kernels do start earlier on the
# reserved SMs but crawl there (200 us instead of 15), and everything downstream waits for the image banks anyway.
_IMGBANK_RESERVE = int(os.environ.get("MGNNS_IMGBANK_RESERVE_SMS", "0"))
_DW_CHUNKS = int(os.environ.get("MGNNS_DW_CHUNKS", "1"))   # measured 1 / 2 / 4 / 8 chunks: 7.34 / 7.45 / 7.49 / 7.87 ms per step
_sm_counts = {}


def _imgbank_ctas():
    """CTA cap for the image-bank kernels: all SMs on a single-stream run, SMs - reserve inside the stream graph."""
    if not _concurrent["on"] or _IMGBANK_RESERVE <= 0:
        return 0
    dev = torch.cuda.current_device()
    n = _sm_counts.get(dev)
    if n is None:
        n = _sm_counts[dev] = torch.cuda.get_device_properties(dev).multi_processor_count
    return max(n - _IMGBANK_RESERVE, 1)


_concurrent = {"on": False}


def set_concurrent_streams(on: bool) -> bool:
    """Tell the ops that other chains of the model are running on other streams (Multi_GCN_Multihead_Att.forward with
    branch_streams, and the backward pass of such a forward)."""
    prev, _concurrent["on"] = _concurrent["on"], bool(on)
    return prev


def _imgbank_impl(fmap, weight, bias):
    fmap, weight, bias = _imgbank_check(fmap, weight, bias)
    B, C, P = fmap.shape
    O = weight.shape[0]
    bank = torch.empty((B, P, O), device=fmap.device, dtype=torch.float32)
    pooled = torch.empty((B, C), device=fmap.device, dtype=torch.float32)
    s = _stream()
    prec = _PRECISIONS[_precision]
    if _tc_ok(C, P, O, 304) and prec == 1:
        # 3xTF32: the operand-splitter pass of the tensor-core kernel also produces the 14x14 global max, so the
        # feature map is read from HBM once; the arg-max is only needed by the feature-map gradient and is
        # recomputed there (empty placeholder here)
        ws = torch.empty((2 * O * C,), device=fmap.device, dtype=torch.float32)
        with _timed("imgbank_fwd"):
            _check(_lib.mgnns_imgbank_fwd_tc_capped(fmap.data_ptr(), weight.data_ptr(), bias.data_ptr(), B, C, P, O,
                                                    prec, ws.data_ptr(), pooled.data_ptr(), bank.data_ptr(), _imgbank_ctas(),
                                                    s), "imgbank_fwd_tc")
        return bank, pooled, torch.empty((0,), device=fmap.device, dtype=torch.int32)
    argmax = torch.empty((B, C), device=fmap.device, dtype=torch.int32)
    with _timed("rowmax"):
        _check(_lib.mgnns_rowmax_f32(fmap.data_ptr(), B * C, P, pooled.data_ptr(), argmax.data_ptr(), s), "rowmax")
    # bank_b [P,O] = fmap_b^T [P,C] @ weight^T [C,O] + bias
    if _tc_ok(C, P, O, 304):
        with _timed("imgbank_fwd"):
            _check(_lib.mgnns_imgbank_fwd_tc(fmap.data_ptr(), weight.data_ptr(), bias.data_ptr(), B, C, P, O,
                                             prec, None, None, bank.data_ptr(), s), "imgbank_fwd_tc")
        return bank, pooled, argmax
    for b0 in range(0, B, 65535):
        nb = min(65535, B - b0)
        with _timed("imgbank_fwd"):
            gemm_raw(1, 1, P, O, C, (fmap, b0 * C * P), P, C * P, weight, C, 0, (bank, b0 * P * O), O, P * O,
                     batch=nb, bias=bias)
    return bank, pooled, argmax


def _imgbank_fake(fmap, weight, bias):
    B, C = fmap.shape[0], fmap.shape[1]
    P = fmap.numel() // (B * C)
    fused = _PRECISIONS[_precision] == 1 and _tc_ok(C, P, weight.shape[0], 304)
    return (fmap.new_empty((B, P, weight.shape[0])), fmap.new_empty((B, C)),
            fmap.new_empty((0,) if fused else (B, C), dtype=torch.int32))


_LIB.impl("imgbank", _imgbank_impl, "CUDA")
torch.library.register_fake("mgnns::imgbank", _imgbank_fake)


def _imgbank_setup(ctx, inputs, output):
    fmap, weight, bias = inputs
    ctx.save_for_backward(fmap, weight, output[2])
    ctx.fshape = fmap.shape
    ctx.leaves = (weight, bias)            # the parameter objects themselves (deferred weight gradients)


def _pick_reduce(batch, target=32):
    r = min(target, batch)
    while batch % r:
        r -= 1
    return r


def _imgbank_backward(ctx, g_bank, g_pooled, g_argmax):
    fmap, weight, argmax = ctx.saved_tensors
    fmap3, weight, _ = _imgbank_check(fmap, weight, weight.new_empty((weight.shape[0],)))
    B, C, P = fmap3.shape
    O = weight.shape[0]
    g_f = g_w = g_b = None
    if g_bank is not None:
        g_bank = _f32c(g_bank, "grad_bank")
    need_w, need_b = ctx.needs_input_grad[1] and g_bank is not None, ctx.needs_input_grad[2] and g_bank is not None

    def weight_grads(max_ctas=0):
        gw = gb = None
        if need_w:
            # gW [O,C] = sum_b gbank_b^T [O,P] @ fmap_b^T [P,C]
            gw = torch.zeros_like(weight)
            r = _pick_reduce(B)
            if _tc_ok(C, P, O, 320):
                # Optional (MGNNS_DW_CHUNKS > 1): cut the product into sample chunks (it accumulates into gw with atomics
                # anyway) so that kernels of other streams waiting for an SM get in at the chunk boundaries — the LSTM
                # weight gradients and the embedding backward otherwise run as a 0.45 ms tail after the last
                # image-bank kernel.  Measured: the per-launch cost of the chunks outweighs the overlap (default 1).
                chunks = min(_DW_CHUNKS, B) if _concurrent["on"] else 1
                with _timed("imgbank_dw"):
                    for ci in range(chunks):
                        b0, b1 = B * ci // chunks, B * (ci + 1) // chunks
                        _check(_lib.mgnns_imgbank_dw_tc_capped(fmap3[b0:b1].data_ptr(), g_bank[b0:b1].data_ptr(), b1 - b0,
                                                               C, P, O, _PRECISIONS[_precision], gw.data_ptr(),
                                                               max_ctas or _imgbank_ctas(), _stream()),
                               "imgbank_dw_tc")
            else:
                with _timed("imgbank_dw"):
                    gemm_raw(1, 1, O, C, P, g_bank, O, P * O, fmap3, P, C * P, gw, C, 0, batch=B, reduce=r, accumulate=1)
        if need_b:
            gb = colsum(g_bank.reshape(B * P, O))
        return [gw, gb]

    if (need_w or need_b) and _defer_state["enabled"] and _HEAVY_PLAN and all(p.is_leaf for p in ctx.leaves):
        # not needed before the optimizer: leaves the backward chain (and is gated behind the LSTM recurrence)
        _run_deferred(weight_grads, (g_bank, fmap3), ctx.leaves, heavy=True)
    elif need_w or need_b:
        g_w, g_b = weight_grads()
    if ctx.needs_input_grad[0]:
        g_f = torch.empty_like(fmap3) if g_bank is not None else torch.zeros_like(fmap3)
        if g_bank is not None:
            # gF_b [C,P] = weight^T [C,O] @ gbank_b^T [O,P]
            for b0 in range(0, B, 65535):
                nb = min(65535, B - b0)
                gemm_raw(1, 1, C, P, O, weight, C, 0, (g_bank, b0 * P * O), O, P * O, (g_f, b0 * C * P), P, C * P,
                         batch=nb)
        if g_pooled is not None:
            g_pooled = _f32c(g_pooled, "grad_pooled")
            if argmax.numel() == 0:
                # fused forward: recompute the first-index arg-max from the feature map (same kernel as the
                # unfused path; only reached when the trunks are trained)
                argmax = torch.empty((B, C), device=fmap3.device, dtype=torch.int32)
                scratch = torch.empty((B, C), device=fmap3.device, dtype=torch.float32)
                _check(_lib.mgnns_rowmax_f32(fmap3.data_ptr(), B * C, P, scratch.data_ptr(), argmax.data_ptr(), _stream()),
                       "rowmax")
            _check(_lib.mgnns_rowmax_bwd_f32(g_pooled.data_ptr(), argmax.data_ptr(), B * C, P, g_f.data_ptr(),
                                             _stream()), "rowmax_bwd")
        g_f = g_f.reshape(ctx.fshape)
    return g_f, g_w, g_b


torch.library.register_autograd("mgnns::imgbank", _imgbank_backward, setup_context=_imgbank_setup)


# ----------------------------------------------------------------------------- PMI counting (no autograd)
def pmi_count(tokens: torch.Tensor, V: int, window: int, pad_id: int, min_count: int, row_range=None):
    """tokens int32 [D,L] on CUDA -> (rowptr int32[V+1], col int32[nnz], cnt int32[nnz], word_count int64[V]).

    No dense [V,V] table: the (centre -> target) pairs are bucketed by centre row and each row is counted in shared
    memory (csrc/pmi_sparse.cu); memory is 12 bytes per emitted pair (35 M pairs = 0.4 GB for 200k TumEmo-shaped
    documents, whatever V is).  `row_range=(lo, hi)` counts only centres in [lo, hi) — disjoint ranges on different
    ranks partition the work with no reduction.  Two host syncs (total pairs, kept cells) size the buffers.
    """
    _need_cuda(tokens)
    if tokens.dtype != torch.int32 or tokens.dim() != 2:
        raise RuntimeError("mgnns pmi_count: tokens must be int32 [D,L]")
    tokens = tokens.contiguous()
    Dn, L = tokens.shape
    dev = tokens.device
    lo, hi = (0, V) if row_range is None else (int(row_range[0]), int(row_range[1]))
    if not (0 <= lo <= hi <= V):
        raise RuntimeError("mgnns pmi_count: bad row_range")
    s = _stream()
    i64 = torch.int64
    row_emit = torch.empty((V,), device=dev, dtype=i64)
    wc = torch.empty((V,), device=dev, dtype=i64)
    _check(_lib.mgnns_pmi_row_emissions(tokens.data_ptr(), Dn, L, V, window, pad_id, lo, hi, row_emit.data_ptr(),
                                        wc.data_ptr(), s), "pmi_row_emissions")
    row_start = torch.empty((V + 1,), device=dev, dtype=i64)
    _check(_lib.mgnns_exclusive_scan_i64(row_emit.data_ptr(), row_start.data_ptr(), V, s), "scan_i64")
    total, biggest = (int(v) for v in torch.stack([row_start[-1], row_emit.max()]).tolist())     # host sync 1
    if biggest >= 2 ** 31:
        raise RuntimeError("mgnns pmi_count: a centre word has %d pairs; per-cell counts are int32 (< 2^31)" % biggest)
    targets = torch.empty((max(total, 1),), device=dev, dtype=torch.int32)
    cursor = torch.empty((V,), device=dev, dtype=i64)
    _check(_lib.mgnns_pmi_scatter_targets(tokens.data_ptr(), Dn, L, V, window, pad_id, lo, hi, row_start.data_ptr(),
                                          cursor.data_ptr(), targets.data_ptr(), s), "pmi_scatter_targets")
    tmp_col = torch.empty((max(total, 1),), device=dev, dtype=torch.int32)
    tmp_cnt = torch.empty((max(total, 1),), device=dev, dtype=torch.int32)
    nnz_row = torch.empty((V,), device=dev, dtype=torch.int32)
    with _timed("pmi_row_reduce"):
        _check(_lib.mgnns_pmi_row_reduce(targets.data_ptr(), row_start.data_ptr(), V, min_count, tmp_col.data_ptr(),
                                         tmp_cnt.data_ptr(), nnz_row.data_ptr(), s), "pmi_row_reduce")
    rowptr = torch.empty((V + 1,), device=dev, dtype=torch.int32)
    _check(_lib.mgnns_exclusive_scan_i32(nnz_row.data_ptr(), rowptr.data_ptr(), V, s), "scan")
    nnz = int(rowptr[-1].item())                                                                # host sync 2
    col = torch.empty((max(nnz, 1),), device=dev, dtype=torch.int32)[:nnz]
    cnt = torch.empty((max(nnz, 1),), device=dev, dtype=torch.int32)[:nnz]
    _check(_lib.mgnns_pmi_compact(tmp_col.data_ptr(), tmp_cnt.data_ptr(), row_start.data_ptr(), rowptr.data_ptr(), V,
                                  col.data_ptr(), cnt.data_ptr(), s), "pmi_compact")
    pmi_count.last_pairs = total
    return rowptr, col, cnt, wc


def pmi_count_dense(tokens: torch.Tensor, V: int, window: int, pad_id: int, min_count: int):
    """Same result through a dense int32 [V,V] table in HBM (1.6 GB at V=20k) and an ordered compaction: the
    round-1 path, kept for small vocabularies and as an independent cross-check of pmi_count (tests)."""
    _need_cuda(tokens)
    if tokens.dtype != torch.int32 or tokens.dim() != 2:
        raise RuntimeError("mgnns pmi_count: tokens must be int32 [D,L]")
    tokens = tokens.contiguous()
    Dn, L = tokens.shape
    dev = tokens.device
    if V * V * 4 > 120 * (1 << 30):
        raise RuntimeError("mgnns pmi_count_dense: V=%d needs a %.0f GB count table" % (V, V * V * 4 / 2**30))
    if Dn * L * 2 * max(window, 1) >= 2 ** 31:
        raise RuntimeError("mgnns pmi_count_dense: the corpus could overflow an int32 cell; use pmi_count")
    pair = torch.zeros((V, V), device=dev, dtype=torch.int32)
    wc = torch.zeros((V,), device=dev, dtype=torch.int64)
    s = _stream()
    _check(_lib.mgnns_pmi_count(tokens.data_ptr(), Dn, L, V, window, pad_id, pair.data_ptr(), wc.data_ptr(), s),
           "pmi_count")
    nnz_row = torch.empty((V,), device=dev, dtype=torch.int32)
    rowptr = torch.empty((V + 1,), device=dev, dtype=torch.int32)
    _check(_lib.mgnns_count_row_nnz_i32(pair.data_ptr(), V, V, min_count, nnz_row.data_ptr(), s), "count_row_nnz")
    _check(_lib.mgnns_exclusive_scan_i32(nnz_row.data_ptr(), rowptr.data_ptr(), V, s), "scan")
    nnz = int(rowptr[-1].item())
    col = torch.empty((max(nnz, 1),), device=dev, dtype=torch.int32)[:nnz]
    cnt = torch.empty((max(nnz, 1),), device=dev, dtype=torch.int32)[:nnz]
    _check(_lib.mgnns_count_fill_csr_i32(pair.data_ptr(), V, V, min_count, rowptr.data_ptr(), col.data_ptr(),
                                         cnt.data_ptr(), s), "count_fill_csr")
    return rowptr, col, cnt, wc


# ----------------------------------------------------------------------------- loop bookkeeping (no autograd)
def confusion_count(scores: torch.Tensor, target: torch.Tensor, conf: torch.Tensor, pred_out: Optional[torch.Tensor] = None):
    """conf[target[b], argmax_c scores[b,c]] += 1 on the device (conf int32 [C,C], accumulated); optional pred_out
    int64 [B].  Replaces the per-batch `.cpu()` + sklearn calls of the reference engine (engine:829-838)."""
    _need_cuda(scores, target, conf, pred_out)
    if scores.dim() != 2 or scores.dtype != torch.float32 or scores.stride(1) != 1:
        raise RuntimeError("mgnns confusion_count: scores must be float32 [B,C] with unit column stride")
    B, C = scores.shape
    if target.dtype != torch.int64 or target.numel() != B or not target.is_contiguous():
        raise RuntimeError("mgnns confusion_count: target must be contiguous int64 [B]")
    if conf.dtype != torch.int32 or conf.numel() != C * C or not conf.is_contiguous():
        raise RuntimeError("mgnns confusion_count: conf must be contiguous int32 [C,C]")
    if pred_out is not None and (pred_out.dtype != torch.int64 or pred_out.numel() != B or not pred_out.is_contiguous()):
        raise RuntimeError("mgnns confusion_count: pred_out must be contiguous int64 [B]")
    _check(_lib.mgnns_confusion_count(scores.data_ptr(), scores.stride(0), target.data_ptr(), B, C, conf.data_ptr(),
                                      _ptr(pred_out), _stream()), "confusion_count")
    return conf


def label_cooccurrence(labels: torch.Tensor, lens: torch.Tensor, C: int, nums: torch.Tensor, adj: torch.Tensor):
    """nums[j] += images containing label j; adj[a,b] += images containing both (a != b) — int64, accumulated.
    labels int32 [n_images, max_len], lens int32 [n_images].  (ref: utils/util.py:336-357)"""
    _need_cuda(labels, lens, nums, adj)
    if labels.dtype != torch.int32 or labels.dim() != 2 or not labels.is_contiguous():
        raise RuntimeError("mgnns label_cooccurrence: labels must be contiguous int32 [n_images, max_len]")
    if lens.dtype != torch.int32 or lens.numel() != labels.shape[0] or not lens.is_contiguous():
        raise RuntimeError("mgnns label_cooccurrence: lens must be contiguous int32 [n_images]")
    if nums.dtype != torch.int64 or nums.numel() != C or adj.dtype != torch.int64 or adj.numel() != C * C \
            or not nums.is_contiguous() or not adj.is_contiguous():
        raise RuntimeError("mgnns label_cooccurrence: nums / adj must be contiguous int64 [C] / [C,C]")
    _check(_lib.mgnns_label_cooccurrence(labels.data_ptr(), lens.data_ptr(), labels.shape[0], labels.shape[1], C,
                                         nums.data_ptr(), adj.data_ptr(), _stream()), "label_cooccurrence")
    return nums, adj


# ----------------------------------------------------------------------------- packed bi-LSTM recurrence
LSTM_TILE = 8


class LstmPlan:
    """Host-side schedule for one batch of variable-length sequences (built from the CPU lengths the
    reference already keeps on the host, engine:804 / model:376 — no device sync).

    Device tensors have a fixed `capacity` (compact rows, >= sum of lengths) so that the same plan object
    can be refreshed in place with `update_()` between replays of a captured CUDA graph:
      tok_idx  int64 [capacity]  compact row -> b*L + t           (padding rows -> 0)
      flat_idx int64 [capacity]  compact row -> b*L + t           (padding rows -> B*L, a dummy bank row)
      last_idx / first_idx int64 [B]  compact row of each sequence's last / first token
      offsets int32 [B+1], lens int32 [B], tiles int32 [ceil(B/8)*8]
    """

    def __init__(self, lens_cpu: torch.Tensor, L: int, device, capacity=None):
        self.B, self.L, self.device = int(lens_cpu.shape[0]), int(L), device
        self.n_tiles = (self.B + LSTM_TILE - 1) // LSTM_TILE
        packed, idx, n = self._host_arrays(lens_cpu, capacity)
        self.capacity = int(idx.shape[0] - 2 * self.B) // 2
        self.N = n
        pin = torch.cuda.is_available()
        self._host_plan = torch.from_numpy(packed).pin_memory() if pin else torch.from_numpy(packed)
        self._host_idx = torch.from_numpy(idx).pin_memory() if pin else torch.from_numpy(idx)
        self.plan = self._host_plan.to(device, non_blocking=True)
        self.idx = self._host_idx.to(device, non_blocking=True)
        B, cap = self.B, self.capacity
        self.tok_idx = self.idx[:cap]
        self.flat_idx = self.idx[cap:2 * cap]
        self.last_idx = self.idx[2 * cap:2 * cap + B]
        self.first_idx = self.idx[2 * cap + B:]
        self.offsets = self.plan[:B + 1]
        self.lens = self.plan[B + 1:2 * B + 1]
        self.tiles = self.plan[2 * B + 1:]

    def _host_arrays(self, lens_cpu, capacity):
        import numpy as np
        B, L = self.B, self.L
        lens = np.clip(lens_cpu.detach().to('cpu', torch.int64).numpy(), 0, L)
        offsets = np.zeros(B + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        n = int(offsets[-1])
        cap = n if capacity is None else int(capacity)
        if cap < n:
            raise RuntimeError("LstmPlan: %d valid tokens exceed the plan capacity %d" % (n, cap))
        order = np.argsort(-lens, kind='stable')
        order = order[lens[order] > 0]
        tiles = np.full(self.n_tiles * LSTM_TILE, -1, dtype=np.int64)
        tiles[:order.shape[0]] = order
        rows = np.repeat(np.arange(B, dtype=np.int64), lens)
        flat = rows * L + (np.arange(n, dtype=np.int64) - offsets[rows])
        tok = np.zeros(cap, dtype=np.int64)
        tok[:n] = flat
        flat_pad = np.full(cap, B * L, dtype=np.int64)
        flat_pad[:n] = flat
        hi = max(n - 1, 0)
        last = np.clip(offsets[1:] - 1, 0, hi)
        first = np.clip(offsets[:-1], 0, hi)
        packed = np.concatenate([offsets, lens, tiles]).astype(np.int32)
        idx = np.concatenate([tok, flat_pad, last, first])
        return packed, idx, n

    def update_(self, lens_cpu: torch.Tensor):
        """Refresh the device tensors in place for a new batch (same B, L and capacity).

        Stream contract: the two H2D copies are enqueued on the CURRENT stream from pinned staging buffers.  The
        staging buffers are rewritten by the host on the next call, so this call first waits (host-side) for the
        event recorded after the previous call's copies — a caller that runs `update_(k); replay(); update_(k+1)`
        without any other synchronisation is safe.  Kernels that read the plan must be ordered after the stream
        this was called on (GraphedTrainStep.update_lengths documents the same)."""
        packed, idx, n = self._host_arrays(lens_cpu, self.capacity)
        self.N = n
        ev = self.__dict__.get('_staged')
        if ev is not None:
            ev.synchronize()                 # the previous update's H2D has finished reading the staging buffers
        self._host_plan.copy_(torch.from_numpy(packed))
        self._host_idx.copy_(torch.from_numpy(idx))
        self.plan.copy_(self._host_plan, non_blocking=True)
        self.idx.copy_(self._host_idx, non_blocking=True)
        if self.plan.is_cuda:
            self._staged = torch.cuda.Event()
            self._staged.record(torch.cuda.current_stream(self.plan.device))
        return self


_LIB.define("lstm_rec(Tensor g, Tensor whh_f, Tensor whh_r, Tensor offsets, Tensor lens, Tensor tiles, int n_tiles) "
            "-> (Tensor, Tensor, Tensor, Tensor)")


def _lstm_check(g, whh_f, whh_r):
    _need_cuda(g, whh_f, whh_r)
    g = _f32c(g, "g")
    whh_f = _f32c(whh_f, "whh_f")
    whh_r = _f32c(whh_r, "whh_r")
    H = whh_f.shape[1]
    if whh_f.shape != (4 * H, H) or whh_r.shape != (4 * H, H) or g.dim() != 2 or g.shape[1] != 8 * H:
        raise RuntimeError("mgnns::lstm_rec: expected g [N,8H], whh_* [4H,H]")
    return g, whh_f, whh_r, H


def _lstm_impl(g, whh_f, whh_r, offsets, lens, tiles, n_tiles):
    g, whh_f, whh_r, H = _lstm_check(g, whh_f, whh_r)
    N = g.shape[0]
    dev = g.device
    s = _stream()
    wt4 = torch.empty((2, H, H, 4), device=dev, dtype=torch.float32)
    _check(_lib.mgnns_lstm_prep_whh(whh_f.data_ptr(), wt4[0].data_ptr(), H, s), "lstm_prep_whh")
    _check(_lib.mgnns_lstm_prep_whh(whh_r.data_ptr(), wt4[1].data_ptr(), H, s), "lstm_prep_whh")
    # rows beyond the valid tokens (plan capacity padding) are never written by the kernel: keep them zero so
    # that the dense GEMMs downstream (next layer, weight gradients) see exact zeros there
    y = torch.zeros((N, 2 * H), device=dev, dtype=torch.float32)
    gates = torch.empty((N, 2, 4, H), device=dev, dtype=torch.float32)
    csave = torch.empty((N, 2, H), device=dev, dtype=torch.float32)
    hprev = torch.zeros((N, 2, H), device=dev, dtype=torch.float32)
    with _timed("lstm_rec_fwd"):
        _check(_lib.mgnns_lstm_rec_fwd(offsets.data_ptr(), lens.data_ptr(), tiles.data_ptr(), n_tiles, H, g.data_ptr(),
                                       wt4[0].data_ptr(), wt4[1].data_ptr(), y.data_ptr(), gates.data_ptr(),
                                       csave.data_ptr(), hprev.data_ptr(), s), "lstm_rec_fwd")
    return y, gates, csave, hprev


def _lstm_fake(g, whh_f, whh_r, offsets, lens, tiles, n_tiles):
    H = whh_f.shape[1]
    N = g.shape[0]
    return g.new_empty((N, 2 * H)), g.new_empty((N, 2, 4, H)), g.new_empty((N, 2, H)), g.new_empty((N, 2, H))


_LIB.impl("lstm_rec", _lstm_impl, "CUDA")
torch.library.register_fake("mgnns::lstm_rec", _lstm_fake)


def _lstm_setup(ctx, inputs, output):
    g, whh_f, whh_r, offsets, lens, tiles, n_tiles = inputs
    ctx.save_for_backward(whh_f, whh_r, offsets, lens, tiles, output[1], output[2], output[3])
    ctx.n_tiles = n_tiles
    ctx.whh_leaves = (whh_f, whh_r)        # the parameter objects themselves (deferred weight gradients)


def _lstm_backward(ctx, gy, ggates, gc, ghp):
    whh_f, whh_r, offsets, lens, tiles, gates, csave, hprev = ctx.saved_tensors
    whh_f = _f32c(whh_f, "whh_f")
    whh_r = _f32c(whh_r, "whh_r")
    H = whh_f.shape[1]
    gy = _f32c(gy, "grad_y")
    N = gy.shape[0]
    dG = torch.zeros((N, 8 * H), device=gy.device, dtype=torch.float32)
    if _LSTM_BWD_AFTER_IMAGE and _defer_state["enabled"] and not _defer_state["gates"]:
        # Optional (measured, no gain): hold the first (top-layer) recurrence of the backward pass until the image-query
        # stacks' backward has delivered the image banks' gradients (the "ready" events of the queued heavy jobs;
        # autograd runs those nodes before this one), so that its 256 whole-SM CTAs do not land in the middle of that
        # chain of small kernels.
        cur = torch.cuda.current_stream(gy.device)
        for job in _defer_state["heavy"]:
            cur.wait_event(job[3])
    _drop_gate(gy.device)                  # gated heavy jobs (image-bank weight gradients) start just behind this launch
    with _timed("lstm_rec_bwd"):
        _check(_lib.mgnns_lstm_rec_bwd(offsets.data_ptr(), lens.data_ptr(), tiles.data_ptr(), ctx.n_tiles, H,
                                       gy.data_ptr(), gates.data_ptr(), csave.data_ptr(), whh_f.data_ptr(),
                                       whh_r.data_ptr(), dG.data_ptr(), _stream()), "lstm_rec_bwd")
    need_f, need_r = ctx.needs_input_grad[1], ctx.needs_input_grad[2]

    def weight_grads():
        g_f_ = g_r_ = None
        if need_f and need_r and _PRECISIONS[_precision] is not None and N >= _TC_MIN_ROWS and (2 * H) % 4 == 0:
            # one tensor-core product dG^T . [Hprev_fwd | Hprev_rev] -> [8H, 2H]; the two diagonal blocks are the
            # gradients (the reverse half of Hprev alone starts at a non-16-byte-aligned address, which TMA rejects;
            # the off-diagonal flops are cheaper than two CUDA-core products)
            full = _mm_impl(dG, hprev.view(N, 2 * H), None, True, False, ACT_NONE, 0.0)
            return [full[:4 * H, :H].contiguous(), full[4 * H:, H:].contiguous()]
        if need_f:
            g_f_ = _mm_impl(dG[:, :4 * H], hprev[:, 0, :], None, True, False, ACT_NONE, 0.0)      # dWhh = dG^T . Hprev
        if need_r:
            g_r_ = _mm_impl(dG[:, 4 * H:], hprev[:, 1, :], None, True, False, ACT_NONE, 0.0)
        return [g_f_, g_r_]

    g_f = g_r = None
    if _defer_state["enabled"] and ctx.whh_leaves[0].is_leaf and ctx.whh_leaves[1].is_leaf:
        _run_deferred(weight_grads, (dG, hprev), ctx.whh_leaves)
    else:
        g_f, g_r = weight_grads()
    return (dG if ctx.needs_input_grad[0] else None), g_f, g_r, None, None, None, None


torch.library.register_autograd("mgnns::lstm_rec", _lstm_backward, setup_context=_lstm_setup)


class _LstmInputProjection(torch.autograd.Function):
    """G[N, 8H] = x . [W_ih_fwd; W_ih_rev]^T + (b_ih + b_hh) for both directions in one many-row product
    (ref: the input half of nn.LSTM, model:179-184,:378).  A Function over the leaf parameters (not cat + mm) so
    that backward can hand the four weight/bias gradients to the deferred stream and return None for them."""

    @staticmethod
    def forward(ctx, x, w_f, w_r, bih_f, bhh_f, bih_r, bhh_r):
        W = torch.cat([w_f, w_r], 0)
        bias = None if bih_f is None else torch.cat([bih_f + bhh_f, bih_r + bhh_r], 0)
        ctx.save_for_backward(x, W)
        ctx.leaves = (w_f, w_r, bih_f, bhh_f, bih_r, bhh_r)
        return _mm_impl(x, W, bias, False, True, ACT_NONE, 0.0)

    @staticmethod
    def backward(ctx, dG):
        x, W = ctx.saved_tensors
        w_f, w_r, bih_f, bhh_f, bih_r, bhh_r = ctx.leaves
        dG = _f32c(dG, "grad")
        H4 = W.shape[0] // 2
        gx = _mm_impl(dG, W, None, False, False, ACT_NONE, 0.0) if ctx.needs_input_grad[0] else None
        need = ctx.needs_input_grad

        def weight_grads():
            gf = _mm_impl(dG[:, :H4], x, None, True, False, ACT_NONE, 0.0) if need[1] else None
            gr = _mm_impl(dG[:, H4:], x, None, True, False, ACT_NONE, 0.0) if need[2] else None
            out = [gf, gr, None, None, None, None]
            if bih_f is not None:
                bf, br = colsum(dG[:, :H4]), colsum(dG[:, H4:])
                out[2:] = [bf if need[3] else None, bf.clone() if need[4] else None,
                           br if need[5] else None, br.clone() if need[6] else None]
            return out

        leaves_ok = all(p is None or p.is_leaf for p in ctx.leaves)
        if _defer_state["enabled"] and leaves_ok:
            _run_deferred(weight_grads, (dG, x), ctx.leaves)
            return (gx, None, None, None, None, None, None)
        return (gx, *weight_grads())


class _PadTextBank(torch.autograd.Function):
    """bank[b, t, :] = y[offsets[b] + t, :] for t < lens[b], zero rows after (ref: pad_packed_sequence(total_length),
    model:386-390) in one pass over the bank; backward gathers the valid rows."""

    @staticmethod
    def forward(ctx, y, offsets, lens, B, L):
        y = _f32c(y, "y")
        F = y.shape[1]
        bank = torch.empty((B, L, F), device=y.device, dtype=torch.float32)
        _check(_lib.mgnns_pad_rows_fwd(y.data_ptr(), offsets.data_ptr(), lens.data_ptr(), B, L, F, bank.data_ptr(), _stream()),
               "pad_rows_fwd")
        ctx.save_for_backward(offsets, lens)
        ctx.dims = (y.shape[0], B, L, F)
        return bank

    @staticmethod
    def backward(ctx, g):
        offsets, lens = ctx.saved_tensors
        N, B, L, F = ctx.dims
        g = _f32c(g, "grad_bank")
        gy = torch.zeros((N, F), device=g.device, dtype=torch.float32)
        _check(_lib.mgnns_pad_rows_bwd(g.data_ptr(), offsets.data_ptr(), lens.data_ptr(), B, L, F, gy.data_ptr(), _stream()),
               "pad_rows_bwd")
        return gy, None, None, None, None


def pad_text_bank(y, plan, B, L):
    """Compact LSTM output [capacity, F] -> zero-padded bank [B, L, F] (F % 4 == 0)."""
    _need_cuda(y)
    if y.shape[1] % 4 != 0:
        raise RuntimeError("mgnns pad_text_bank: feature width must be a multiple of 4")
    return _PadTextBank.apply(y, plan.offsets, plan.lens, int(B), int(L))


class _EmbeddingRows(torch.autograd.Function):
    """weight[tokens] (ref: self.embedding(text), model:377) with the table's gradient as ONE scatter-add launch
    (rows equal to padding_idx receive no gradient, as nn.Embedding(padding_idx=...) does)."""

    @staticmethod
    def forward(ctx, weight, tokens, padding_idx):
        ctx.save_for_backward(tokens)
        ctx.meta = (weight.shape, -1 if padding_idx is None else int(padding_idx))
        return weight.index_select(0, tokens)

    @staticmethod
    def backward(ctx, g):
        (tokens,) = ctx.saved_tensors
        (V, E), pad = ctx.meta
        g = _f32c(g, "grad")
        gw = torch.zeros((V, E), device=g.device, dtype=torch.float32)
        _check(_lib.mgnns_embedding_bwd(tokens.data_ptr(), g.data_ptr(), tokens.numel(), E, pad, V, gw.data_ptr(), _stream()),
               "embedding_bwd")
        return gw, None, None


def embedding_rows(weight, tokens, padding_idx=None):
    _need_cuda(weight, tokens)
    if weight.dtype != torch.float32 or weight.dim() != 2 or weight.shape[1] % 4 != 0 or not weight.is_contiguous():
        raise RuntimeError("mgnns embedding_rows: weight must be contiguous float32 [V, E] with E % 4 == 0")
    if tokens.dtype != torch.int64 or tokens.dim() != 1 or not tokens.is_contiguous():
        raise RuntimeError("mgnns embedding_rows: tokens must be contiguous int64 [N]")
    return _EmbeddingRows.apply(weight, tokens, padding_idx)


def packed_bilstm(lstm: torch.nn.LSTM, x_compact: torch.Tensor, plan: LstmPlan, training: bool, after_first_projection=None):
    """Multi-layer bidirectional LSTM over compacted tokens with nn.LSTM's parameters (same names, same
    math as torch's packed-sequence path, inter-layer dropout included).  Returns [N, 2H]."""
    if not lstm.bidirectional or not lstm.batch_first or lstm.proj_size != 0:
        raise NotImplementedError("mgnns_b200 packed_bilstm: bidirectional, batch_first, no projection only")
    x = x_compact
    for layer in range(lstm.num_layers):
        sfx = "_l%d" % layer
        P = lambda name: getattr(lstm, name + sfx)                     # noqa: E731
        R = lambda name: getattr(lstm, name + sfx + "_reverse")        # noqa: E731
        if lstm.bias:
            biases = (P("bias_ih"), P("bias_hh"), R("bias_ih"), R("bias_hh"))
        else:
            biases = (None, None, None, None)
        _need_cuda(x)
        x = _f32c(x, "x")
        g = _LstmInputProjection.apply(x, P("weight_ih"), R("weight_ih"), *biases)                  # [N, 8H]
        if after_first_projection is not None:
            # called once per layer, just before its (latency-critical, SM-hungry) recurrence is enqueued — e.g. to
            # record an event that releases a persistent kernel of another stream behind this launch
            after_first_projection()
        y = torch.ops.mgnns.lstm_rec(g, P("weight_hh"), R("weight_hh"), plan.offsets, plan.lens, plan.tiles, plan.n_tiles)[0]
        if layer + 1 < lstm.num_layers and training and lstm.dropout > 0:
            y = torch.nn.functional.dropout(y, lstm.dropout, True)
        x = y
    return x
