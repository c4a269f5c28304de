"""Per-phase cycle counts of the LSTM forward cluster kernel (CTA (0,0) = the longest tile): product, barrier, gate phase,
barrier, exchange wait.  python scripts/lstm_phases.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from mgnns_b200 import ops, _abi
dev = torch.device('cuda', 0)
lib = _abi.lib
lib.mgnns_lstm_debug_buffer.restype = ctypes.c_int
lib.mgnns_lstm_debug_buffer.argtypes = [ctypes.c_void_p]
buf = torch.zeros(16, dtype=torch.int64, device=dev)
lstm = torch.nn.LSTM(300, 150, num_layers=2, batch_first=True, bidirectional=True).to(dev)
hb = bench.host_batch(512, 0)
plan = ops.LstmPlan(hb['lens'], 100, dev, None)
x = torch.randn(plan.capacity, 300, device=dev)
with torch.no_grad():
    for _ in range(3):
        ops.packed_bilstm(lstm, x, plan, False)
    torch.cuda.synchronize()
    assert lib.mgnns_lstm_debug_buffer(buf.data_ptr()) == 0
    ops.packed_bilstm(lstm, x, plan, False)
    torch.cuda.synchronize()
    lib.mgnns_lstm_debug_buffer(None)
b = buf.cpu().tolist()
for name, o in (('thread 0', 0), ('thread 608', 8)):
    n = max(b[o + 5], 1)
    print(name, 'steps', b[o + 5], ' cycles/step: product %.0f  sync1 %.0f  gates %.0f  sync2 %.0f  exchange %.0f  total %.0f'
          % tuple([b[o + i] / n for i in range(5)] + [sum(b[o:o + 5]) / n]))
