"""Per-phase timing of the table-free PMI count (CUDA events around each C-ABI call): python scripts/pmi_phases.py [V] [docs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mgnns_b200 import ops, synth
from mgnns_b200._abi import lib, check
dev = torch.device('cuda', 0)
for V, docs in ((20154, 200000), (50000, 200000)):
    text, lens, _ = synth.make_texts(docs, V, 100, seed=7)
    tok = text.to(torch.int32).to(dev)
    s = torch.cuda.current_stream().cuda_stream
    i64, i32 = torch.int64, torch.int32
    row_emit = torch.empty(V, device=dev, dtype=i64); wc = torch.empty(V, device=dev, dtype=i64)
    row_start = torch.empty(V + 1, device=dev, dtype=i64); cursor = torch.empty(V, device=dev, dtype=i64)
    def ev():
        e = torch.cuda.Event(enable_timing=True); e.record(); return e
    for rep in range(3):
        t = [ev()]
        check(lib.mgnns_pmi_row_emissions(tok.data_ptr(), docs, 100, V, 6, 0, 0, V, row_emit.data_ptr(), wc.data_ptr(), s)); t.append(ev())
        check(lib.mgnns_exclusive_scan_i64(row_emit.data_ptr(), row_start.data_ptr(), V, s)); t.append(ev())
        total = int(row_start[-1].item())
        targets = torch.empty(total, device=dev, dtype=i32); tmp_col = torch.empty(total, device=dev, dtype=i32); tmp_cnt = torch.empty(total, device=dev, dtype=i32)
        nnz_row = torch.empty(V, device=dev, dtype=i32); rowptr = torch.empty(V + 1, device=dev, dtype=i32)
        t.append(ev())
        check(lib.mgnns_pmi_scatter_targets(tok.data_ptr(), docs, 100, V, 6, 0, 0, V, row_start.data_ptr(), cursor.data_ptr(), targets.data_ptr(), s)); t.append(ev())
        check(lib.mgnns_pmi_row_reduce(targets.data_ptr(), row_start.data_ptr(), V, 2, tmp_col.data_ptr(), tmp_cnt.data_ptr(), nnz_row.data_ptr(), s)); t.append(ev())
        check(lib.mgnns_exclusive_scan_i32(nnz_row.data_ptr(), rowptr.data_ptr(), V, s)); t.append(ev())
        nnz = int(rowptr[-1].item())
        col = torch.empty(nnz, device=dev, dtype=i32); cnt = torch.empty(nnz, device=dev, dtype=i32)
        t.append(ev())
        check(lib.mgnns_pmi_compact(tmp_col.data_ptr(), tmp_cnt.data_ptr(), row_start.data_ptr(), rowptr.data_ptr(), V, col.data_ptr(), cnt.data_ptr(), s)); t.append(ev())
        torch.cuda.synchronize()
    names = ['emissions', 'scan64', 'sync+alloc', 'scatter', 'row_reduce', 'scan32', 'sync+alloc', 'compact']
    print("V=%d docs=%d pairs=%d kept=%d max row=%d :: " % (V, docs, total, nnz, int(row_emit.max())) +
          "  ".join("%s %.3f" % (n, a.elapsed_time(b)) for n, a, b in zip(names, t[:-1], t[1:])) + "  total %.3f ms" % t[0].elapsed_time(t[-1]))
