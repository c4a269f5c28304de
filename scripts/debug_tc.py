"""Debug probe for the tcgen05 image-bank kernels (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgnns_b200 import _abi
lib, check = _abi.lib, _abi.check
dev = torch.device('cuda', 0)
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)

def fwd(F, W, bias, prec=0):
    B, C, P = F.shape
    O = W.shape[0]
    out = torch.full((B, P, O), -777.0, device=dev)
    check(lib.mgnns_imgbank_fwd_tc(F.data_ptr(), W.data_ptr(), bias.data_ptr(), B, C, P, O, prec, out.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream), 'fwd')
    torch.cuda.synchronize()
    return out

def dw(F, G, prec=0):
    B, C, P = F.shape
    O = G.shape[2]
    out = torch.zeros((O, C), device=dev)
    check(lib.mgnns_imgbank_dw_tc(F.data_ptr(), G.data_ptr(), B, C, P, O, prec, out.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream), 'dw')
    torch.cuda.synchronize()
    return out

for C in (32, 64):
    B, P, O = 1, 196, 300
    zero = torch.zeros(O, device=dev)
    F = torch.ones(B, C, P, device=dev); W = torch.ones(O, C, device=dev)
    o = fwd(F, W, zero)
    print('C=%d ones: expect %d; got min %.3f max %.3f; corner' % (C, C, o.min().item(), o.max().item()), o[0, :3, :4].flatten().tolist())
    F = torch.arange(P, device=dev).float().view(1, 1, P).expand(B, C, P).contiguous()
    o = fwd(F, W, zero)
    ref = torch.einsum('bcp,oc->bpo', F, W)
    print(' F=p: max err', (o - ref).abs().max().item(), 'rows', o[0, [0, 1, 2, 31, 32, 33, 127, 128, 195], 0].tolist())
    F = torch.ones(B, C, P, device=dev); W = torch.arange(O, device=dev).float().view(O, 1).expand(O, C).contiguous()
    o = fwd(F, W, zero); ref = torch.einsum('bcp,oc->bpo', F, W)
    print(' W=o: max err', (o - ref).abs().max().item(), 'cols', o[0, 0, [0, 1, 2, 7, 8, 9, 159, 160, 161, 299]].tolist())
    F = torch.randn(B, C, P, device=dev); W = torch.randn(O, C, device=dev)
    o = fwd(F, W, zero); ref = torch.einsum('bcp,oc->bpo', F.double(), W.double()).float()
    print(' rand tf32: max err', (o - ref).abs().max().item(), ' 3x:', (fwd(F, W, zero, 1) - ref).abs().max().item())
B, C, P, O = 3, 256, 196, 300
F = torch.randn(B, C, P, device=dev); G = torch.randn(B, P, O, device=dev)
ref = torch.einsum('bpo,bcp->oc', G.double(), F.double()).float()
print('dw tf32 max err', (dw(F, G) - ref).abs().max().item(), '3x', (dw(F, G, 1) - ref).abs().max().item(), 'ref max', ref.abs().max().item())
