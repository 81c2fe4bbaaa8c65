"""K7 (request) sweep on one B200: 1000 queries against a 50k-signature index (BASELINE configs[2]),
queries/s for a few grid sizes (GSB_K7_CTAS) at ef_search 5000 and 1600."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gsearch_b200 as g
from gsearch_b200.comm import DeviceBuffer

S, n, nq = 18000, int(os.environ.get("DB", "50000")), 1000
db = g.synth.signatures(n, S)
q, _ = g.synth.queries(nq, db)
d_db = DeviceBuffer(db.nbytes); d_db.upload(db)
idx = g.Hnsw(g.HnswParams(max_nb_conn=128, ef=1600), S, np.uint64)
t0 = time.perf_counter()
idx.insert_device(d_db.ptr, np.arange(n, dtype=np.uint64))
print("build s", time.perf_counter() - t0, flush=True)
for ef in (5000, 1600):
    for ctas in sys.argv[1:] or ["0"]:
        if ctas != "0": os.environ["GSB_K7_CTAS"] = ctas
        else: os.environ.pop("GSB_K7_CTAS", None)
        idx.search_raw(q, 50, ef)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); o, c, ne = idx.search_raw(q, 50, ef); ts.append(time.perf_counter() - t0)
        t = min(ts)
        gb = ne.mean() * S * 8 * nq / t / 1e9
        print(f"ef={ef} ctas={ctas:>4s} q/s={nq / t:9.1f} (host call) evals/q={ne.mean():.0f} GB/s={gb:7.1f} frac={gb / 6451.5:.3f}", flush=True)
