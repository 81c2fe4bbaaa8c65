"""K6 (batched DistHamming) alone: nq queries x n candidates, device resident, CUDA events.
Algorithmic bytes = nq * n * S * sizeof(Sig) when n * S * sizeof(Sig) exceeds L2 (every query
streams the whole candidate matrix once)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gsearch_b200 as g  # noqa: E402
from gsearch_b200 import _lib  # noqa: E402

S, n, nq = 18000, int(sys.argv[1]) if len(sys.argv) > 1 else 8192, int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda", 0)
cands = torch.randint(1, 2**40, (n, S), dtype=torch.int64, device=dev)
q = cands[:nq].clone()
q[:, ::3] += 1
out = torch.empty(nq, n, dtype=torch.float32, device=dev)
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream


def run():
    _lib.check(L.gsb_hamming_matrix_dev(C.c_void_p(q.data_ptr()), nq, C.c_void_p(cands.data_ptr()), n, S,
                                        g.SIG_U64, C.c_void_p(out.data_ptr()), C.c_void_p(st)))


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 5
e0.record()
for _ in range(K):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
byts = nq * n * S * 8
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
d = out.cpu().numpy()
assert abs(d[0, 0] - 6000 / 18000) < 1e-6 and d[0, 1] == 1.0
print(json.dumps({"kernel": "k6_hamming_matrix", "nq": nq, "n": n, "S": S, "ms": ms,
                  "achieved_gbs": byts / 1e9 / (ms / 1e3), "peak_gbs": peak,
                  "frac": byts / 1e9 / (ms / 1e3) / peak, "evals_per_s": nq * n / (ms / 1e3)}))
