import ctypes as C, sys, os
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import gsearch_b200 as g
from gsearch_b200 import _lib
sys.argv=['bench.py']
import bench
L=_lib.lib()
dev=torch.device('cuda',0)
S,n,nq=18000,8192,296
base=bench.tree_signatures(torch,n,S,dev,1234)
gen=torch.Generator(device=dev); gen.manual_seed(99)
pick=torch.randint(0,n,(nq,),device=dev,generator=gen)
fresh=torch.randint(1,2**40,(nq,S),dtype=torch.int64,device=dev,generator=gen)
q=torch.where(torch.rand(nq,S,device=dev,generator=gen)<0.9,base[pick],fresh).cpu().numpy().view(np.uint64)
idx=g.Hnsw(g.HnswParams(max_nb_conn=128,ef=1600),S,np.uint64)
idx.insert_device(base.data_ptr(),np.arange(n,dtype=np.uint64))
out=(C.c_ulonglong*4)()
L.gsb_debug_k7_prof(out)
import time
idx.search_raw(q,50,1600); L.gsb_debug_k7_prof(out)
t0=time.perf_counter(); o,c,ne=idx.search_raw(q,50,1600); dt=time.perf_counter()-t0
L.gsb_debug_k7_prof(out)
tot=out[0]+out[1]+out[2]
print("search_layer calls",out[3],"eval %.3f heap %.3f gather %.3f"%(out[0]/tot,out[1]/tot,out[2]/tot),"cycles/query",tot/nq, "wall ms",dt*1e3, "evals/query", ne.mean())
