import sys, os, time, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import gsearch_b200 as g
from gsearch_b200 import _lib
_a = sys.argv; sys.argv = ['bench.py']
import bench
sys.argv = _a
dev = torch.device('cuda', 0)
S, n = 18000, 16384
base = bench.tree_signatures(torch, n, S, dev, 1234)
torch.cuda.synchronize()
idx = g.Hnsw(g.HnswParams(max_nb_conn=128, ef=1600), S, np.uint64)
out = (C.c_ulonglong * 8)()
L = _lib.lib()
L.gsb_debug_k8_prof(out)
t0 = time.perf_counter()
idx.insert_device(base.data_ptr(), np.arange(n, dtype=np.uint64))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
L.gsb_debug_k8_prof(out)
tot = sum(out[i] for i in range(4))
print("inserts/s", round(n / dt), "layers", out[4], "search %.3f mates+rebuild %.3f select %.3f sort %.3f" % tuple(out[i] / tot for i in range(4)),
      "cycles/layer", tot / max(out[4], 1), "mean selected", out[5] / max(out[4], 1))
