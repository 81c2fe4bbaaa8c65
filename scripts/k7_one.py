"""one K7 launch at BASELINE configs[2]: EF env = ef_search.
  ncu:           ncu --set full --clock-control none -k regex:k7_hnsw -c 1 -o out python scripts/k7_one.py
  phase timing:  make -C gsearch_b200/csrc clean all EXTRA=-DGSB_K7_PROF ; python scripts/k7_one.py
                 (clock64 per phase of the search loop, printed by a few CTAs; rebuild without EXTRA afterwards)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gsearch_b200 as g
from gsearch_b200.comm import DeviceBuffer
S, n, nq = 18000, int(os.environ.get("DB", "50000")), 1000
db = g.synth.signatures(n, S)
q, _ = g.synth.queries(nq, db)
d_db = DeviceBuffer(db.nbytes); d_db.upload(db)
idx = g.Hnsw(g.HnswParams(max_nb_conn=128, ef=1600), S, np.uint64)
idx.insert_device(d_db.ptr, np.arange(n, dtype=np.uint64))
o, c, ne = idx.search_raw(q, 50, int(os.environ.get("EF", "5000")))
print("evals/q", ne.mean())
