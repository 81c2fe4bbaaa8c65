import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import gsearch_b200 as g
_a = sys.argv; sys.argv = ['bench.py']
import bench
sys.argv = _a
dev = torch.device('cuda', 0)
S, n = 18000, 32768
base = bench.tree_signatures(torch, n, S, dev, 1234)
torch.cuda.synchronize()
q = base[::111].cpu().numpy().view(np.uint64)[:148]
for w in (148, 296, 444, 148, 296):
    idx = g.Hnsw(g.HnswParams(max_nb_conn=128, ef=1600), S, np.uint64)
    idx.set_wave_max(w)
    t0 = time.perf_counter()
    idx.insert_device(base.data_ptr(), np.arange(n, dtype=np.uint64))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out, cnt, ne = idx.search_raw(q, 10, 1600)
    self_hit = float((out["d_id"][:, 0] == np.arange(0, n, 111)[:148]).mean())
    print("wave", w, "inserts/s", round(n / dt), "self-hit", self_hit, "evals/query", float(ne.mean()))
    idx.close()
