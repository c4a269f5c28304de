"""Host-side mirror of the reference's Python interface for the hot path.

Module map (reference -> here):
  models/Multi_GCN_Multihead_att.py -> multi_gcn.py
  models/Text_GCN.py                -> text_gcn.py
  models/submodules.py, moudles.py  -> layers.py
  utils/pmi.py                      -> pmi.py
  utils/util.py (gen_A, gen_adj)    -> graph_util.py
"""
