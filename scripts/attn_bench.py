"""Single-query attention core timing (CUDA events), image-bank (L=196) and masked text-bank (L=100) shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mgnns_b200 import ops, synth
dev = torch.device('cuda', 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.manual_seed(0)
for L, masked in ((196, False), (100, True)):
    us, banks = [], []
    for i in range(6):      # rotate over > L2 worth of banks
        us.append((torch.randn(B, 4, 300, device=dev) * 0.1).requires_grad_())
        banks.append(torch.randn(B, L, 300, device=dev).requires_grad_())
    mask, rows = None, B * L
    if masked:
        _, lens, m = synth.make_texts(B, 20154, L, seed=1)
        mask = m.to(dev)
        rows = int(lens.clamp(max=L).sum())
    g = torch.randn(B, 4, 300, device=dev)
    def fwd(i):
        return torch.ops.mgnns.attn_q1(us[i % 6], banks[i % 6], mask, 0.088, 0.1, 123)[0]
    for i in range(3):
        fwd(i).backward(g)
    torch.cuda.synchronize()
    ops.KernelTimers.reset(['attn_q1_fwd', 'attn_q1_bwd'])
    for i in range(12):
        fwd(i).backward(g)
    torch.cuda.synchronize()
    f, _ = ops.KernelTimers.mean_ms('attn_q1_fwd')
    b, _ = ops.KernelTimers.mean_ms('attn_q1_bwd')
    ops.KernelTimers.reset([])
    byt = rows * 1200
    print("L=%d masked=%s B=%d live rows=%d: fwd %.1f us (%.0f GB/s)  bwd %.1f us (%.0f GB/s, bank read + dbank write over all L)"
          % (L, masked, B, rows, f * 1e3, byt / f / 1e6, b * 1e3, (byt + B * L * 1200) / b / 1e6))
