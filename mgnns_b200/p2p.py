"""Gradient all-reduce over NVLink peer memory (mgnns_allreduce_p2p_f32, csrc/p2p_allreduce.cu).

One process per GPU (torch.distributed is only the rendezvous: it carries the 64-byte CUDA IPC handles).  Each rank
allocates its flat gradient buffer and a small flag block with mgnns_p2p_alloc (plain cudaMalloc, so CUDA IPC can export
it), maps the other ranks' allocations, and from then on the all-reduce is ONE kernel launch on the caller's stream —
which, unlike a process-group collective, a multi-stream CUDA-graph capture records like any other kernel
(ref: the training step engine/Multi_GCN_Multihead_Att_engine.py:847-851).
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _abi

_lib = _abi.lib
_check = _abi.check


class _RawCuda:
    """A cudaMalloc'd range exposed through __cuda_array_interface__ so torch can alias it as a tensor."""

    def __init__(self, ptr, numel, typestr):
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _alloc(nbytes):
    out = ctypes.c_void_p()
    _check(_lib.mgnns_p2p_alloc(int(nbytes), ctypes.byref(out)), "p2p_alloc")
    return int(out.value)


def _export(ptr):
    h = ctypes.create_string_buffer(64)
    _check(_lib.mgnns_p2p_export(ptr, h), "p2p_export")
    return h.raw


def _import(handle):
    out = ctypes.c_void_p()
    _check(_lib.mgnns_p2p_import(handle, ctypes.byref(out)), "p2p_import")
    return int(out.value)


class PeerMemoryUnavailable(RuntimeError):
    """Raised on EVERY rank when any rank cannot allocate, export or map the peer buffers (the decision is agreed through
    the process group, so callers can fall back to NCCL consistently)."""


class PeerAllReduce:
    """`flat` (float32 [numel], numel % 4 == 0) lives in peer-mapped memory; all_reduce_(scale) makes it
    scale * (sum over ranks) on every rank, bit-identically, with one kernel on the current stream.  Every rank must
    call all_reduce_ the same number of times (the cross-GPU barriers count epochs on the device)."""

    def __init__(self, numel, device, group=None, ctas=None):
        if not dist.is_initialized():
            raise RuntimeError("PeerAllReduce needs an initialised torch.distributed process group (rendezvous only)")
        if numel % 4:
            raise ValueError("PeerAllReduce: numel must be a multiple of 4 floats")
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.numel = int(numel)
        self.ctas = int(ctas if ctas is not None else os.environ.get('MGNNS_P2P_CTAS', '96'))
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            ok, err, mine = True, None, None
            try:
                self._buf = _alloc(4 * max(self.numel, 4))
                self._flags = _alloc(_lib.mgnns_p2p_flag_bytes())
                mine = (_export(self._buf), _export(self._flags))
            except Exception as exc:                                   # agreed on below: nobody is left in a collective
                ok, err = False, str(exc)
            if os.environ.get('MGNNS_P2P_FAIL_RANK') == str(self.rank):  # test hook for the fallback path
                ok, err, mine = False, 'forced failure (MGNNS_P2P_FAIL_RANK)', None
            handles = [None] * self.world
            dist.all_gather_object(handles, (ok, err, mine), group=group)   # also a barrier: every allocation is zeroed by now
            bad = [(q, h[1]) for q, h in enumerate(handles) if not h[0]]
            if bad:
                raise PeerMemoryUnavailable("rank %d could not set up its peer-mapped buffers: %s" % bad[0])
            bufs, flags = [], []
            self._imported = []
            try:
                for q, (_, _, (hb, hf)) in enumerate(handles):
                    if q == self.rank:
                        bufs.append(self._buf)
                        flags.append(self._flags)
                    else:
                        pb, pf = _import(hb), _import(hf)
                        self._imported += [pb, pf]
                        bufs.append(pb)
                        flags.append(pf)
            except Exception as exc:
                ok, err = False, str(exc)
            agreed = torch.tensor([1 if ok else 0], device=self.device, dtype=torch.int32)
            dist.all_reduce(agreed, op=dist.ReduceOp.MIN, group=group)     # nobody launches before everybody has mapped
            if int(agreed.item()) == 0:
                raise PeerMemoryUnavailable("a rank could not map its peers' buffers through CUDA IPC%s"
                                            % (": " + err if err else ""))
        self._bufs = (ctypes.c_uint64 * self.world)(*bufs)
        self._flagv = (ctypes.c_uint64 * self.world)(*flags)
        self._holder = _RawCuda(self._buf, self.numel, "<f4")
        self.flat = torch.as_tensor(self._holder, device=self.device)
        assert self.flat.data_ptr() == self._buf and self.flat.dtype == torch.float32

    def all_reduce_(self, scale=1.0):
        s = torch.cuda.current_stream(self.device).cuda_stream
        _check(_lib.mgnns_allreduce_p2p_f32(self._bufs, self._flagv, self.rank, self.world, self.numel, float(scale),
                                            self.ctas, s), "allreduce_p2p")
        return self.flat

    def check(self):
        """Raise if a cross-GPU barrier ever timed out on this rank (synchronises the current stream)."""
        rc = _lib.mgnns_p2p_error(self._flags, torch.cuda.current_stream(self.device).cuda_stream)
        if rc != 0:
            raise RuntimeError("PeerAllReduce: %s" % ("a cross-GPU barrier timed out (a peer did not launch its all-reduce)"
                                                      if rc > 0 else "CUDA error while reading the status word"))

    def close(self):
        torch.cuda.synchronize(self.device)
        for p in getattr(self, "_imported", []):
            _lib.mgnns_p2p_close(p)
        self._imported = []
