#!/usr/bin/env python
"""bench.py -- genomes/sec sketched (`tohnsw`) and queries/sec (`request`) on N B200s of one node.

Headline line (BASELINE.json metric, configs[1]): ProbMinHash3a signatures of synthetic 5 Mbp
genomes, k=21, s=18000 (Sig = u64).  A step = one pass of the hot path (FASTA bytes -> signatures)
over `--reps` sub-batches of `--batch` genomes per GPU, rotating through `--pool` DISTINCT
sub-batches (each ~0.5 GB of FASTA, larger than L2: nothing is served from cache between calls).
`value` is measured with the FASTA resident in HBM (CUDA events, max over ranks); `e2e` goes through
the public host-pointer C-ABI call (gsb_sketch_fasta_batch) with the FASTA in pinned host memory:
H2D + kernels + D2H of the signatures inside the timed region, over the same number of steps.
With N > 1 every rank sketches its own genomes (weak scaling) and the finished signatures are
all-gathered over NCCL inside the timed region of BOTH numbers, as `tohnsw` needs them on every GPU
before HNSW insertion; after the timed region every rank re-sketches genomes owned by OTHER ranks
and compares them with the gathered rows (multi-GPU parity, asserted).

The same JSON line carries a "request" sub-record (BASELINE configs[2] at full size): a
50 000-signature index built on device (s=18000, n=128, ef=1600), 1 000 queries, at the
reference's hard-coded ef_search = 5000 (src/bin/gsearch.rs:893) and at 1600; queries/s with the
queries resident in HBM (`value`) and through gsb_index_search_batch with host buffers (`e2e`),
K7 roofline from the distance evaluations the kernel reports, CPU arm beside it.  With N > 1 the
index is replicated (built in parallel on every GPU from the same signatures), queries shard by
rank, answers are all-gathered.

`--impl reference` times the reference's CPU implementation of the same two paths: the C
restatement under oracle/ (the Rust reference cannot be built in this image -- no cargo, crates
not vendored), one genome / one query per host thread as the reference does
(src/dna/dnasketch.rs:325, parallel_search), on all host cores, each step a bounded sample.  That
process maps no product code: the synthetic data comes from datagen/libgsb_synth.so.

Other workloads: `--aa --algo optdens --kmer 7 --sketch 12000` (configs[3]); `--workload request`
prints the request record alone; `--no-request` skips it.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "genomes/sec sketched (tohnsw sketch phase)"
UNIT = "genomes/s"
EF_SEARCH_REFERENCE = 5000  # src/bin/gsearch.rs:893


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--batch", type=int, default=96, help="genomes per library call (sub-batch) per GPU")
    ap.add_argument("--pool", type=int, default=6, help="distinct sub-batches rotated through")
    ap.add_argument("--reps", type=int, default=24, help="sub-batches per step (a step = reps x batch genomes per GPU)")
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--kmer", type=int, default=21)
    ap.add_argument("--sketch", type=int, default=18000)
    ap.add_argument("--algo", default="prob", choices=["prob", "optdens", "super"])
    ap.add_argument("--aa", action="store_true", help="proteomes (BASELINE configs[3]: --aa --algo optdens "
                                                      "--kmer 7 --sketch 12000): 4500 proteins x ~333 aa per file")
    ap.add_argument("--nprot", type=int, default=4500)
    ap.add_argument("--cpu-sample", type=int, default=0, help="genomes per CPU step (0 = 2 per host thread)")
    ap.add_argument("--workload", default="both", choices=["both", "sketch", "request"])
    ap.add_argument("--no-request", action="store_true")
    ap.add_argument("--db", type=int, default=50000, help="request: signatures in the index")
    ap.add_argument("--queries", type=int, default=1000, help="request: queries per step")
    ap.add_argument("--nbng", type=int, default=128, help="request: max_nb_connection (-n of tohnsw)")
    ap.add_argument("--ef", type=int, default=1600, help="request: ef_construction (--ef of tohnsw)")
    ap.add_argument("--ef-search", type=int, nargs="*", default=[EF_SEARCH_REFERENCE, 1600])
    ap.add_argument("--knbn", type=int, default=50)
    ap.add_argument("--cpu-queries", type=int, default=0, help="request: queries per CPU step (0 = 4 per host thread)")
    ap.add_argument("--cpu-ef-construction", type=int, default=320,
                    help="reference arm: ef_construction of the CPU-built index (bounded build; search unchanged)")
    ap.add_argument("--min-window", type=float, default=2.0, help="request: seconds per timed window")
    a = ap.parse_args()
    if a.no_request and a.workload == "both":
        a.workload = "sketch"
    if a.aa and a.workload == "both":
        a.workload = "sketch"
    return a


ALGO_ID = {"prob": 0, "super": 1, "optdens": 2}


def sketch_config(a):
    """identical in both arms (the driver compares it)"""
    if a.aa:
        wl = (f"configs[3]: {a.algo} sketch of synthetic proteomes ({a.nprot} proteins x ~333 aa), k={a.kmer} "
              f"s={a.sketch} --aa")
    else:
        wl = (f"configs[1]: ProbMinHash3a sketch of synthetic {a.genome_len / 1e6:g} Mbp genomes, k={a.kmer} "
              f"s={a.sketch} --algo {a.algo} (distinct genomes of the 10k-genome set)")
    return {"workload": wl, "kmer": a.kmer, "sketch_size": a.sketch, "algo": a.algo,
            "unit_bytes": "one FASTA file per genome, 80-column lines", "l2": "inputs larger than L2 (no flush needed)"}


def request_config(a, ef_search):
    return {"workload": f"configs[2]: {a.queries} queries vs {a.db}-signature HNSW (s={a.sketch} n={a.nbng} "
                        f"ef={a.ef}), ef_search={ef_search}, knbn={a.knbn}",
            "l2": f"index signatures {a.db * a.sketch * 8 / 1e9:.2f} GB, larger than L2"}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def synth():
    """datagen/libgsb_synth.so through its ctypes wrapper; imports no product library"""
    from gsearch_b200 import synth as s
    return s


def gen_files(a, first, n, threads):
    """n synthetic FASTA files (bytes objects), generated on `threads` host threads"""
    S = synth()
    fn = (lambda i: S.aa_proteome(first + i, a.nprot)) if a.aa else (lambda i: S.dna_genome(first + i, a.genome_len))
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        return list(ex.map(fn, range(n)))


def concat(files):
    offs = np.zeros(len(files) + 1, dtype=np.uint64)
    np.cumsum([len(f) for f in files], out=offs[1:])
    return np.frombuffer(b"".join(files), dtype=np.uint8), offs


# --------------------------------------------------------------------------- reference arm
def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    cores = host_threads()
    line = None
    if a.workload in ("both", "sketch"):
        algo = ALGO_ID[a.algo]
        sample = a.cpu_sample or 2 * cores
        buf, offs = concat(gen_files(a, 0, sample, cores))
        dt_ = 1 if a.aa else 0
        for _ in range(a.warmup):
            O.sketch_buffer(buf, offs, a.kmer, a.sketch, algo, dt_, False, 0, nthreads=cores)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            O.sketch_buffer(buf, offs, a.kmer, a.sketch, algo, dt_, False, 0, nthreads=cores)
        dt = time.perf_counter() - t0
        value = sample * a.steps / dt
        what = f"{a.nprot}-protein proteomes" if a.aa else f"{a.genome_len} bp genomes"
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64" if (algo == 0 and not a.aa) else "f32",
            "data": "synthetic", "config": sketch_config(a), "units_per_step": sample,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} x {what} per step x {a.steps} steps, one genome per host thread, "
                                       f"C restatement of the reference CPU path (oracle/)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
    if a.workload in ("both", "request"):
        req = reference_request(a, O, cores)
        if line is None:
            line = dict(req)
            line.update({"impl": "reference", "n_gpus": a.gpus, "higher_is_better": True, "scaling": "weak",
                         "vs_baseline": None, "dtype": "u64", "data": "synthetic", "gpu_launches": 0})
        else:
            line["request"] = req
    print(json.dumps(line), flush=True)


def reference_request(a, O, cores):
    """CPU arm of `request`: signatures from the same generator, index built by the oracle's wave
    insertion on all host cores (ef_construction bounded so that the build ends within about a
    minute; M and the data are the GPU arm's), then parallel_search, one query per thread."""
    S = synth()
    t0 = time.perf_counter()
    db = S.signatures(a.db, a.sketch, np.uint64, seed=1234)
    queries, _ = S.queries(a.queries, db, seed=99, noise=0.1)
    t_gen = time.perf_counter() - t0
    h = O.Hnsw(a.nbng, a.cpu_ef_construction, a.sketch, np.uint64)
    t0 = time.perf_counter()
    h.insert_waves(db, np.arange(a.db, dtype=np.uint64), 296, nthreads=cores)
    t_build = time.perf_counter() - t0
    del db
    sample = min(a.queries, a.cpu_queries or 4 * cores)
    recs = {}
    for ef in a.ef_search:
        for _ in range(min(a.warmup, 1)):
            h.search(queries[:sample], a.knbn, ef, nthreads=cores)
        steps = max(1, min(a.steps, 5))
        t0 = time.perf_counter()
        for _ in range(steps):
            out, cnt, neval = h.search(queries[:sample], a.knbn, ef, nthreads=cores)
        dt = time.perf_counter() - t0
        qps = sample * steps / dt
        recs[ef] = {
            "metric": "queries/sec (request)", "value": qps, "unit": "queries/s", "steps": steps,
            "warmup": min(a.warmup, 1), "ms_per_step": 1e3 * dt / steps, "config": request_config(a, ef),
            "units_per_step": sample, "mean_evaluations_per_query": float(neval.mean()),
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} queries per step x {steps} steps, one query per host thread, oracle "
                                       f"search (oracle/hnsw.c) on an index the oracle built"},
        }
    head = dict(recs[a.ef_search[0]])
    head["build"] = {"seconds": t_build, "genomes_per_s_inserted": a.db / t_build,
                     "ef_construction_used": a.cpu_ef_construction,
                     "note": "CPU build bounded by a smaller ef_construction than the GPU arm's; the timed search "
                             "uses the same M, data, ef_search and knbn", "datagen_seconds": t_gen}
    for ef in a.ef_search[1:]:
        head[f"ef_search_{ef}"] = recs[ef]
    return head


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={device}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
            time.sleep(0.3)  # the sampler needs a moment to start: short windows came back empty
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, power, reasons = [], [], [], set()
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
                power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                               c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "power_w_max": float(max(power)), "samples": len(sm)}
        return out


def measured_traffic(key):
    """DRAM bytes from the committed ncu capture (profiles/r2_traffic.json), or None"""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", name)))[key]
            d["source"] = f"profiles/{name}"
            return d
        except Exception:
            continue
    return None


def k7_traffic(a, ef, world):
    """ncu DRAM bytes of one K7 launch, if the committed capture is of this very configuration"""
    t = measured_traffic("k7_hnsw_search_ef%d" % ef)
    if t and world == 1 and t.get("queries") == a.queries and t.get("db_signatures") == a.db:
        return t["dram_bytes_per_launch"]
    return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- own arm
def dist_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def run_own(a):
    import torch
    import torch.distributed as dist
    import gsearch_b200 as g

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the library has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    line = None
    if a.workload in ("both", "sketch"):
        line = own_sketch(a, torch, dist, g, rank, world, local, dev)
    if a.workload in ("both", "request"):
        req = own_request(a, torch, dist, g, rank, world, local, dev)
        if rank == 0:
            if line is None:
                line = req
            else:
                line["request"] = req
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def own_sketch(a, torch, dist, g, rank, world, local, dev):
    S_ = synth()
    algo = ALGO_ID[a.algo]
    params = g.SeqSketcherParams(a.kmer, a.sketch, algo, g.DATA_AA if a.aa else g.DATA_DNA)
    sk = g.Sketcher(params, device=local)
    elem = sk.elem_size
    B, S, P, R = a.batch, a.sketch, max(1, a.pool), max(1, a.reps)
    threads = max(1, host_threads() // max(1, world))

    # ---- P distinct sub-batches of this rank, generated straight into pinned host memory.  Genome
    # index = ((p * world) + rank) * B + i: every (rank, sub-batch, i) is a different genome
    cap1 = S_.max_bytes(a.nprot * (333 + 333 // 2 + 2), a.nprot) if a.aa else S_.max_bytes(a.genome_len, 1)
    h_bytes = torch.empty(cap1 * B * P, dtype=torch.uint8, pin_memory=True)
    base_ptr = h_bytes.data_ptr()

    def gindex(p, r, i):
        return (p * world + r) * B + i

    def gen_one(t):
        p, i = divmod(t, B)
        dst = base_ptr + (p * B + i) * cap1
        if a.aa:
            return S_.aa_proteome_into(gindex(p, rank, i), a.nprot, 333, dst, cap1)
        return S_.dna_genome_into(gindex(p, rank, i), a.genome_len, 1, dst, cap1)

    with ThreadPoolExecutor(max_workers=threads) as ex:
        lens = list(ex.map(gen_one, range(P * B)))
    assert all(n > 0 for n in lens)
    # compact each sub-batch in place (files back to back), offsets per sub-batch
    offs, sub_off, sub_len = [], [], []
    hb = h_bytes.numpy()
    for p in range(P):
        o = np.zeros(B + 1, dtype=np.uint64)
        start = p * B * cap1
        pos = 0
        for i in range(B):
            n = lens[p * B + i]
            src = start + i * cap1
            if src != start + pos:
                hb[start + pos:start + pos + n] = hb[src:src + n]
            pos += n
            o[i + 1] = pos
        offs.append(o)
        sub_off.append(start)
        sub_len.append(pos)
    total_per_sub = float(np.mean(sub_len))
    d_bytes = [torch.empty(sub_len[p] + 64, dtype=torch.uint8, device=dev) for p in range(P)]
    for p in range(P):
        d_bytes[p][:sub_len[p]].copy_(h_bytes[sub_off[p]:sub_off[p] + sub_len[p]], non_blocking=True)
    h_sig = torch.empty(B * S * elem, dtype=torch.uint8, pin_memory=True)
    h_nb = torch.empty(B, dtype=torch.int64, pin_memory=True)
    d_sig = torch.empty(B * S * elem, dtype=torch.uint8, device=dev)
    d_nb = torch.empty(B, dtype=torch.int64, device=dev)
    d_all = torch.empty(world * B * S * elem, dtype=torch.uint8, device=dev) if world > 1 else None
    stream = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()

    def call_resident(p):
        sk.sketch_device(d_bytes[p].data_ptr(), offs[p], B, d_sig.data_ptr(), d_nb.data_ptr(), stream)
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_sig)

    def call_e2e(p):
        sk.sketch_pointers(base_ptr + sub_off[p], offs[p], B, h_sig.data_ptr(), h_nb.data_ptr())
        if world > 1:
            d_sig.copy_(h_sig, non_blocking=True)  # the library returned host signatures: back for the gather
            dist.all_gather_into_tensor(d_all, d_sig)

    def step(fn, s):
        for r in range(R):
            fn((s * R + r) % P)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(3, a.warmup)
    for s in range(W):
        step(call_resident, s)
    barrier()

    # ---- timed region 1: HBM-resident inputs
    sk.enable_timing(True)
    launches0 = sk.launch_count
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for s in range(a.steps):
        step(call_resident, s)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = sk.launch_count - launches0
    ktimes = sk.kernel_times()
    sk.enable_timing(False)
    clk = clocks.stop() if clocks else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    per_step = world * B * R
    value = per_step * a.steps / (ms / 1e3)

    # sanity, outside the timed region: every signature is filled and encoded lengths are right
    last_p = ((a.steps - 1) * R + R - 1) % P
    nb = d_nb.cpu().numpy()
    if a.aa:
        assert (nb > 0.5 * a.nprot * 333).all() and (nb < 2 * a.nprot * 333).all(), nb[:4]
    else:
        assert (nb > 0.99 * a.genome_len - 100).all() and (nb <= a.genome_len).all(), nb[:4]
    sig_last = d_sig.cpu().numpy().view(sk.dtype).reshape(B, S).copy()
    assert (sig_last != 0).mean() > 0.999

    # ---- multi-GPU parity (outside the timed regions): rows of the gathered matrix that OTHER ranks
    # produced must equal what this GPU computes for the same global genomes
    parity = None
    if world > 1:
        gathered = d_all.cpu().numpy().view(sk.dtype).reshape(world, B, S)
        assert np.array_equal(gathered[rank], sig_last), "own shard changed in the all-gather"
        checked = 0
        for r in range(world):
            if r == rank:
                continue
            picks = sorted({0, B // 2, B - 1})
            if a.aa:
                files = [S_.aa_proteome(gindex(last_p, r, i), a.nprot) for i in picks]
            else:
                files = [S_.dna_genome(gindex(last_p, r, i), a.genome_len) for i in picks]
            mine, _ = sk.sketch_files(files)
            for k, i in enumerate(picks):
                assert np.array_equal(mine[k], gathered[r, i]), \
                    f"rank {rank}: gathered signature of genome ({r},{i}) differs from the local recomputation"
                checked += 1
        tchk = torch.tensor([checked], dtype=torch.int64, device=dev)
        dist.all_reduce(tchk)
        parity = {"gathered_rows_recomputed_on_other_ranks": int(tchk.item()), "equal": True}

    # ---- timed region 2, e2e: host pointers through the public C-ABI call (H2D + kernels + D2H
    # every call, + the all-gather at N > 1), same number of steps
    step(call_e2e, 0)  # warm the staging buffers
    barrier()
    t0 = time.perf_counter()
    for s in range(a.steps):
        step(call_e2e, s)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = per_step * a.steps / float(t.item())
    assert np.array_equal(h_sig.numpy().view(sk.dtype).reshape(B, S), sig_last), "e2e and resident paths differ"

    # ---- the H2D link alone, all ranks copying at the same time: what e2e is bounded by on this box
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for r in range(2 * P):
        p = r % P
        d_bytes[p][:sub_len[p]].copy_(h_bytes[sub_off[p]:sub_off[p] + sub_len[p]], non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    tcopy = torch.tensor([c0.elapsed_time(c1) / 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tcopy, op=dist.ReduceOp.MAX)
    h2d_gbs_rank = 2 * sum(sub_len) / 1e9 / float(tcopy.item())

    if rank != 0:
        return None
    peak, peak_src = measured_peak_gbs()
    fasta_per_genome = total_per_sub / B
    alg_bytes_per_genome = fasta_per_genome + S * elem  # L_fasta + S*sizeof(Sig)
    # The kernels of two genome groups and the FASTA packer of the next group run concurrently on
    # three streams, so a single kernel's own duration is not separable inside the pipeline; the span
    # timed here (CUDA events on the launching stream, first group start -> last group end) covers
    # the whole K1 || K2 || K3 pipeline of one library call.
    k2_ms, k2_n = ktimes["k2_scan"]
    traffic = measured_traffic("sketch_prob_k21_s18000") if (algo == 0 and not a.aa) else None
    calls_timed = R * a.steps
    achieved = (alg_bytes_per_genome * B * calls_timed / 1e9) / (k2_ms / 1e3) if k2_ms > 0 else None
    cfg = sketch_config(a)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": W, "ms_per_step": ms / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64" if elem == 8 else ("f32" if algo else "u32"),
        "data": "synthetic", "config": cfg, "units_per_step": per_step,
        "step": {"genomes_per_call_per_gpu": B, "calls_per_step": R, "distinct_sub_batches": P,
                 "fasta_bytes_per_step_per_gpu": int(total_per_sub * R),
                 "collective": "NCCL all_gather_into_tensor of signatures after every call" if world > 1 else "none",
                 "timed_window_s": ms / 1e3},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(total_per_sub * R),
                "d2h_bytes_per_step": int((B * S * elem + B * 8) * R), "steps": a.steps,
                "timed_window_s": float(t.item()),
                "h2d_copy_alone_gbs_per_gpu": h2d_gbs_rank, "h2d_copy_alone_gbs_aggregate": h2d_gbs_rank * world,
                "h2d_bound_genomes_per_s": world * h2d_gbs_rank * 1e9 / fasta_per_genome,
                "frac_of_h2d_bound": e2e_value / (world * h2d_gbs_rank * 1e9 / fasta_per_genome),
                "note": "h2d_copy_alone: every rank copies its pinned FASTA at the same time (max over ranks)"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {
            "kernel": ("prob sketch pipeline (k2p_count 45 % and k2p_partition 36 % of the kernel time, K1 parse + "
                       "pack 17 %; concurrent streams)" if algo == 0 else "k2_optdens"),
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": (achieved / peak) if achieved else None,
            "traffic": (traffic["dram_bytes_per_genome"] * B if traffic else None),
            "traffic_source": (traffic["source"] + ": ncu dram bytes of all sketch kernels per genome x genomes "
                               "per launch") if traffic else None,
            "peak_source": peak_src, "algorithmic_bytes_per_genome": alg_bytes_per_genome,
            "algorithmic_bytes_per_launch": alg_bytes_per_genome * B,
            "avg_launch_ms": (k2_ms / k2_n) if k2_n else None, "launches_timed": k2_n,
            "note": "integer-ALU bound (two SplitMix64 mixes + bijective mix + rolling per k-mer), not HBM bound "
                    "(DESIGN.md); fraction against HBM reported as required; kernel_ms holds the per-family stream "
                    "times (they overlap across streams)",
        },
        "kernel_ms": {k: v[0] for k, v in ktimes.items()},
        "retries": int(sk.retry_count), "fallbacks": int(sk.fallback_count),
    }
    if parity:
        line["multi_gpu_parity"] = parity
    if world == 1:
        line["cpu_baseline"] = cpu_baseline(a)
    return line


def cpu_baseline(a):
    """the oracle (C port of the reference CPU path) on this box's host cores, bounded sample"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    cores = host_threads()
    sample = a.cpu_sample or 2 * cores
    buf, offs = concat(gen_files(a, 0, sample, cores))
    algo = ALGO_ID[a.algo]
    t0 = time.perf_counter()
    reps = 0
    while reps < 1 or time.perf_counter() - t0 < 8.0:
        O.sketch_buffer(buf, offs, a.kmer, a.sketch, algo, 1 if a.aa else 0, False, 0, nthreads=cores)
        reps += 1
    dt = time.perf_counter() - t0
    what = f"{a.nprot}-protein proteomes" if a.aa else f"{a.genome_len} bp genomes"
    return {"value": sample * reps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample} x {what} x {reps} passes, one per thread (oracle/, C restatement)"}


# --------------------------------------------------------------------------- request workload
def own_request(a, torch, dist, g, rank, world, local, dev):
    """queries/sec of `request`.  The index is built on device by gsb_index_insert_batch_dev (timed:
    genomes/s inserted); `value` = queries/s with the queries resident in HBM
    (gsb_index_search_batch_dev), `e2e` = through gsb_index_search_batch with host queries in and
    host neighbours out.  Roofline: every distance evaluation streams one candidate signature
    (S * 8 bytes) from HBM; the kernel reports the evaluations it performed."""
    S_ = synth()
    S, n, nq = a.sketch, a.db, a.queries
    # rank 0 generates the database once; the other GPUs get it over NCCL (replicated index)
    d_db = torch.empty((n, S), dtype=torch.int64, device=dev)
    h_q = None
    t0 = time.perf_counter()
    if rank == 0:
        db = S_.signatures(n, S, np.uint64, seed=1234)
        h_q, _ = S_.queries(nq, db, seed=99, noise=0.1)
        d_db.copy_(torch.from_numpy(db.view(np.int64)))
        db_head = db[:64].copy()
        del db
    t_gen = time.perf_counter() - t0
    d_q_all = torch.empty((nq, S), dtype=torch.int64, device=dev)
    if rank == 0:
        d_q_all.copy_(torch.from_numpy(h_q.view(np.int64)))
    if world > 1:
        dist.broadcast(d_db, 0)
        dist.broadcast(d_q_all, 0)
    torch.cuda.synchronize()
    idx = g.Hnsw(g.HnswParams(max_nb_conn=a.nbng, ef=a.ef), S, np.uint64, device=local)
    comm = None
    if world > 1:
        # HNSW construction sharded over the GPUs (gsb_index_insert_batch_sharded): inside every
        # insertion wave a rank searches / selects for its slice of the points, the selections are
        # all-gathered (NCCL behind the C ABI) and every rank applies the same link updates
        uid_t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid_t.copy_(torch.frombuffer(bytearray(g.comm.unique_id()), dtype=torch.uint8))
        dist.broadcast(uid_t, 0)
        comm = g.comm.Comm(bytes(uid_t.cpu().numpy().tobytes()), world, rank, local)
        idx.set_wave_max(min(1024, 296 * world))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if comm is not None:
        idx.insert_sharded(comm, d_db.data_ptr(), np.arange(n, dtype=np.uint64))
    else:
        idx.insert_device(d_db.data_ptr(), np.arange(n, dtype=np.uint64))
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    build_parity = None
    if world > 1:
        tb = torch.tensor([t_build], dtype=torch.float64, device=dev)
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        t_build = float(tb.item())
        # every replica must hold the same graph: compare a checksum of the exported adjacency
        import hashlib
        gr = idx.export_graph()
        hsh = hashlib.sha256(gr["nbr_index"].tobytes() + gr["nbr_offsets"].tobytes() + gr["levels"].tobytes()).digest()
        hv = torch.tensor([int.from_bytes(hsh[:7], "little")], dtype=torch.int64, device=dev)
        lo, hi = hv.clone(), hv.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert int(lo.item()) == int(hi.item()), "sharded insertion left different graphs on different GPUs"
        build_parity = {"replica_graph_checksums_equal": True, "links": int(len(gr["nbr_index"]))}
        del gr
    # this rank's queries: j mod world == rank (src/dna/dnarequest.rs:353: one query = one task)
    mine = torch.arange(rank, nq, world, device=dev)
    d_q = d_q_all[mine].contiguous()
    nq_loc = d_q.shape[0]
    hq_loc = d_q.cpu().numpy().view(np.uint64)
    per = (nq + world - 1) // world
    item = 24  # sizeof(gsb_neighbour)
    d_out = torch.zeros((per, a.knbn * item), dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros(per, dtype=torch.int32, device=dev)
    d_nev = torch.zeros(per, dtype=torch.int64, device=dev)
    g_out = torch.empty((world * per, a.knbn * item), dtype=torch.uint8, device=dev) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, ef):
        """-> (seconds of the window [max over ranks], calls)"""
        fn(ef)
        barrier()
        t0 = time.perf_counter()
        fn(ef)
        torch.cuda.synchronize()
        one = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(one, op=dist.ReduceOp.MAX)
        calls = max(2, int(np.ceil(a.min_window / max(float(one.item()), 1e-4))))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(calls):
            fn(ef)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        tt = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()), calls

    def search_dev(ef):
        idx.search_device(d_q.data_ptr(), nq_loc, a.knbn, ef, d_out.data_ptr(), d_cnt.data_ptr(), d_nev.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(g_out, d_out)

    host_res = {}
    # host side of the e2e call: queries and answers in pinned memory (the copies run at link speed)
    from gsearch_b200.comm import PinnedBuffer
    hq_c = np.ascontiguousarray(hq_loc)
    pin_q = PinnedBuffer(max(hq_c.nbytes, 16))
    pin_q.array[:hq_c.nbytes] = hq_c.view(np.uint8).reshape(-1)
    pin_out = PinnedBuffer(max(nq_loc * a.knbn * item, 16))
    pin_cnt = PinnedBuffer(max(nq_loc * 4, 16))
    pin_nev = PinnedBuffer(max(nq_loc * 8, 16))

    def search_host(ef):
        idx.search_pointers(pin_q.ptr, nq_loc, a.knbn, ef, pin_out.ptr, pin_cnt.ptr, pin_nev.ptr)
        host_res["r"] = (pin_out.array[:nq_loc * a.knbn * item].view(g.index.NEIGHBOUR_DTYPE).reshape(nq_loc, a.knbn),
                         pin_cnt.array[:nq_loc * 4].view(np.uint32), pin_nev.array[:nq_loc * 8].view(np.uint64))
        if world > 1:  # the answers go back to the device for the gather
            d_out[:nq_loc].copy_(torch.from_numpy(host_res["r"][0].view(np.uint8).reshape(nq_loc, -1)), non_blocking=True)
            dist.all_gather_into_tensor(g_out, d_out)

    peak, peak_src = measured_peak_gbs()
    recs = {}
    clocks = ClockSampler(local) if rank == 0 else None
    for ef in a.ef_search:
        dt, calls = timed(search_dev, ef)
        nev = d_nev[:nq_loc].clone()
        tot_ev = nev.sum().to(torch.float64)
        if world > 1:
            dist.all_reduce(tot_ev)
        qps = nq * calls / dt
        dte, callse = timed(search_host, ef)
        qps_e2e = nq * callse / dte
        out, cnt, neval = host_res["r"]
        # device-resident and host paths must return the same neighbours
        dev_out = d_out[:nq_loc].cpu().numpy().view(g.index.NEIGHBOUR_DTYPE).reshape(nq_loc, a.knbn)
        assert np.array_equal(dev_out["d_id"], out["d_id"]) and np.array_equal(d_cnt[:nq_loc].cpu().numpy(), cnt)
        bytes_step = float(tot_ev.item()) * S * 8
        achieved = bytes_step * calls / 1e9 / dt / world  # per GPU
        recs[ef] = {
            "metric": "queries/sec (request)", "value": qps, "unit": "queries/s", "n_gpus": world, "steps": calls,
            "warmup": 2, "ms_per_step": 1e3 * dt / calls, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
            "config": request_config(a, ef), "units_per_step": nq, "timed_window_s": dt,
            "e2e": {"value": qps_e2e, "unit": "queries/s", "h2d_bytes_per_step": int(nq * S * 8),
                    "d2h_bytes_per_step": int(nq * a.knbn * item + nq * 12), "steps": callse, "timed_window_s": dte,
                    "note": "gsb_index_search_batch: host queries in, host neighbours out"},
            "gpu_launches": calls,
            "roofline": {"kernel": "k7_hnsw_search", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": k7_traffic(a, ef, world), "peak_source": peak_src,
                         "algorithmic_bytes_per_query": bytes_step / nq,
                         "mean_evaluations_per_query": float(tot_ev.item()) / nq,
                         "avg_launch_ms": 1e3 * dt / calls},
        }
        recs[ef]["_host"] = (out.copy(), cnt.copy())  # (views of the pinned buffers: the next call overwrites them)
    clk = clocks.stop() if clocks else None
    # multi-GPU parity: the gathered answers of every query equal a single-GPU search of that query
    parity = None
    if world > 1:
        ef = a.ef_search[-1]
        search_dev(ef)
        torch.cuda.synchronize()
        full = g_out.view(world, per, -1).transpose(0, 1).reshape(world * per, -1)[:nq]
        picks = torch.arange((rank + 1) % world, nq, max(world, nq // 24), device=dev)[:24]
        dq = d_q_all[picks].contiguous()
        o2 = torch.zeros((len(picks), a.knbn * item), dtype=torch.uint8, device=dev)
        c2 = torch.zeros(len(picks), dtype=torch.int32, device=dev)
        idx.search_device(dq.data_ptr(), len(picks), a.knbn, ef, o2.data_ptr(), c2.data_ptr(), 0)
        torch.cuda.synchronize()
        assert torch.equal(o2, full[picks]), f"rank {rank}: gathered answers differ from the local search"
        tchk = torch.tensor([len(picks)], dtype=torch.int64, device=dev)
        dist.all_reduce(tchk)
        parity = {"gathered_answers_recomputed_on_other_ranks": int(tchk.item()), "equal": True}
    if rank != 0:
        return None
    # recall against brute force on a few queries (size-independent sanity inside the bench)
    out, cnt = recs[a.ef_search[0]].pop("_host")
    for ef in a.ef_search[1:]:
        recs[ef].pop("_host")
    nchk = min(8, nq_loc)
    h_db = d_db.cpu().numpy().view(np.uint64)
    d = g.DistHamming().matrix(hq_loc[:nchk], h_db)
    hits = sum(int((out["distance"][i][:cnt[i]] <= np.sort(d[i])[a.knbn - 1]).sum()) for i in range(nchk))
    head = recs[a.ef_search[0]]
    head["clocks"] = clk
    head["build"] = {"genomes_per_s_inserted": n / t_build, "seconds": t_build, "kernel": "k8_hnsw_insert_select",
                     "ef_construction": a.ef, "datagen_seconds": t_gen,
                     "sharding": ("gsb_index_insert_batch_sharded: phase A of every wave split over the GPUs, "
                                  "selections all-gathered, wave_max %d" % min(1024, 296 * world)) if world > 1
                     else "single GPU, wave_max 296"}
    if build_parity:
        head["build"]["multi_gpu_parity"] = build_parity
    head["recall_at_knbn_on_8_queries"] = hits / (nchk * a.knbn)
    if parity:
        head["multi_gpu_parity"] = parity
    for ef in a.ef_search[1:]:
        head[f"ef_search_{ef}"] = recs[ef]
    if world == 1:
        # CPU arm: the oracle's search on the SAME graph, one query per host thread (parallel_search)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import _oracle as O
        cores = host_threads()
        h = O.Hnsw(a.nbng, a.ef, S, np.uint64)
        h.import_graph(h_db, idx.export_graph())
        cq = min(nq, a.cpu_queries or 4 * cores)
        for ef in a.ef_search:
            t0 = time.perf_counter()
            want, wcnt, _ = h.search(hq_loc[:cq], a.knbn, ef, nthreads=cores)
            dtc = time.perf_counter() - t0
            ref_out = (out if ef == a.ef_search[0] else idx.search_raw(hq_loc[:cq], a.knbn, ef)[0])
            assert want["d_id"].tolist() == ref_out["d_id"][:cq].tolist(), "GPU and CPU answers differ on the same graph"
            rec = head if ef == a.ef_search[0] else head[f"ef_search_{ef}"]
            rec["cpu_baseline"] = {"value": cq / dtc, "unit": "queries/s", "cores": cores, "kind": "port",
                                   "sample": f"{cq} queries, one per thread, oracle search on the graph the GPU "
                                             f"built (answers asserted identical)"}
    return head


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)
