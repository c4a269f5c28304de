"""All-reduce microbench: mgnns_allreduce_p2p_f32 (peer memory, CTA sweep) against NCCL on the cfg-4 / cfg-5 gradient payloads.
Run under torchrun on N GPUs: python -m torch.distributed.run --nproc-per-node N scripts/p2p_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from mgnns_b200.p2p import PeerAllReduce

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


for numel in (24_858_240, 63_490_204 // 4 * 4):
    x = torch.randn(numel, device=dev)
    ms = timed(lambda: dist.all_reduce(x))
    moved = 4 * numel * (world - 1) / world
    if rank == 0:
        print("numel %d (%.1f MB)  NCCL %.3f ms  (%.0f GB/s each way per rank)" % (numel, 4 * numel / 1e6, ms, moved / ms / 1e6), flush=True)
    for ctas in (32, 64, 96, 128):
        peer = PeerAllReduce(numel, dev, ctas=ctas)
        peer.flat.copy_(x)
        ms = timed(lambda: peer.all_reduce_(1.0 / world))
        peer.check()
        if rank == 0:
            print("   p2p ctas=%3d  %.3f ms  (%.0f GB/s each way per rank)" % (ctas, ms, moved / ms / 1e6), flush=True)
        peer.close()
        del peer
dist.barrier()
dist.destroy_process_group()
