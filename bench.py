#!/usr/bin/env python
"""bench.py -- genomes/sec sketched (the `tohnsw` sketch phase) on N B200s of one node.

Workload (BASELINE.json configs[1]): ProbMinHash3a signatures of synthetic 5 Mbp genomes,
k=21, s=18000 (Sig = u64).  A step = one pass of the hot path (FASTA bytes -> signatures) over
one batch of `--batch` distinct genomes per GPU; the batch (~0.5 GB of FASTA) is larger than
L2, so nothing is served from cache between steps.  `value` is measured with the batch
resident in HBM (CUDA events, max over ranks); `e2e` goes through the public host-pointer C-ABI
call with the FASTA in pinned host memory (H2D + kernels + D2H of the signatures inside the
timed region; the H2D copy of later genome groups runs under the kernels of earlier ones).
With N > 1 every rank sketches its own genomes (weak scaling) and the finished signatures are
all-gathered over NCCL inside the timed region, as `tohnsw` needs them on every GPU before HNSW
insertion.

Other workloads: `--aa --algo optdens --kmer 7 --sketch 12000` (configs[3]); `--workload request`
(configs[2]: a 50 000-signature HNSW index built on device, 1 000 queries, K7 roofline from the
distance evaluations the kernel reports, CPU arm = oracle search on the same graph).

`--impl reference` times the reference's CPU implementation of the same path: the C
restatement under oracle/ (the Rust reference cannot be built in this image -- no cargo, crates
not vendored), one genome per host thread as the reference does
(src/dna/dnasketch.rs:325), on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "genomes/sec sketched (tohnsw sketch phase)"
UNIT = "genomes/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--batch", type=int, default=96, help="genomes per step per GPU")
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--kmer", type=int, default=21)
    ap.add_argument("--sketch", type=int, default=18000)
    ap.add_argument("--algo", default="prob", choices=["prob", "optdens", "super"])
    ap.add_argument("--aa", action="store_true", help="proteomes (BASELINE configs[3]: --aa --algo optdens "
                                                      "--kmer 7 --sketch 12000): 4500 proteins x ~333 aa per file")
    ap.add_argument("--nprot", type=int, default=4500)
    ap.add_argument("--cpu-sample", type=int, default=32, help="genomes in the CPU baseline sample")
    # secondary workload (BASELINE configs[2] shape): build an HNSW index on device, then search it
    ap.add_argument("--workload", default="sketch", choices=["sketch", "request"])
    ap.add_argument("--db", type=int, default=50000, help="request: signatures in the index")
    ap.add_argument("--queries", type=int, default=1000, help="request: queries per step")
    ap.add_argument("--nbng", type=int, default=128, help="request: max_nb_connection (-n of tohnsw)")
    ap.add_argument("--ef", type=int, default=1600, help="request: ef_construction (--ef of tohnsw)")
    ap.add_argument("--ef-search", type=int, default=1600)
    ap.add_argument("--knbn", type=int, default=50)
    ap.add_argument("--cpu-queries", type=int, default=16, help="request: queries in the CPU baseline sample")
    return ap.parse_args()


ALGO_ID = {"prob": 0, "super": 1, "optdens": 2}


def workload_name(a):
    if a.aa:
        return (f"configs[3]: {a.algo} sketch of synthetic proteomes ({a.nprot} proteins x ~333 aa), k={a.kmer} "
                f"s={a.sketch} --aa; step = batch of {a.batch} distinct proteomes per GPU")
    return (f"configs[1]: ProbMinHash3a sketch of synthetic {a.genome_len / 1e6:g} Mbp genomes, k={a.kmer} "
            f"s={a.sketch} --algo {a.algo}; step = batch of {a.batch} distinct genomes per GPU "
            f"(slice of the 10k-genome set)")


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def gen_batch_numpy(first_index, n, length):
    """n synthetic genomes -> (uint8 array, uint64 offsets); CPU arm only"""
    import gsearch_b200 as g
    files = [g.synth.dna_genome(first_index + i, length) for i in range(n)]
    return g.Sketcher.concat(files)


def gen_batch_aa_numpy(first_index, n, nprot):
    import gsearch_b200 as g
    return g.Sketcher.concat([g.synth.aa_proteome(first_index + i, nprot) for i in range(n)])


# --------------------------------------------------------------------------- reference arm
def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    algo = ALGO_ID[a.algo]
    cores = host_threads()
    sample = max(1, min(a.batch, a.cpu_sample))
    buf, offs = gen_batch_aa_numpy(0, sample, a.nprot) if a.aa else gen_batch_numpy(0, sample, a.genome_len)
    dt_ = 1 if a.aa else 0
    for _ in range(max(1, min(a.warmup, 1))):
        O.sketch_buffer(buf, offs, a.kmer, a.sketch, algo, dt_, False, 0, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        O.sketch_buffer(buf, offs, a.kmer, a.sketch, algo, dt_, False, 0, nthreads=cores)
    dt = time.perf_counter() - t0
    value = sample * a.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(a), "step_sample": f"{sample} genomes per step on the host CPU"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} x {a.genome_len} bp genomes x {a.steps} steps, one genome per "
                                   f"thread, C restatement of the reference CPU path (oracle/)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={device}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
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
    """DRAM bytes from the committed ncu capture (profiles/r1_traffic.json), or None"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))[key]
    except Exception:
        return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- own arm
def run_own(a):
    import torch
    import torch.distributed as dist
    import gsearch_b200 as g

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the library has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    algo = ALGO_ID[a.algo]
    params = g.SeqSketcherParams(a.kmer, a.sketch, algo, g.DATA_AA if a.aa else g.DATA_DNA)
    sk = g.Sketcher(params, device=local)
    elem = sk.elem_size
    B, S = a.batch, a.sketch

    # ---- synthetic batch of this rank, generated straight into pinned host memory
    cap1 = g.synth.max_bytes(a.nprot * (333 + 333 // 2 + 2), a.nprot) if a.aa else g.synth.max_bytes(a.genome_len, 1)
    h_bytes = torch.empty(cap1 * B, dtype=torch.uint8, pin_memory=True)
    offs = np.zeros(B + 1, dtype=np.uint64)
    pos = 0
    for i in range(B):
        if a.aa:
            n = g.synth.aa_proteome_into(rank * B + i, a.nprot, 333, h_bytes.data_ptr() + pos, cap1)
        else:
            n = g.synth.dna_genome_into(rank * B + i, a.genome_len, 1, h_bytes.data_ptr() + pos, cap1)
        assert n > 0
        pos += n
        offs[i + 1] = pos
    total = pos
    h_sig = torch.empty(B * S * elem, dtype=torch.uint8, pin_memory=True)
    h_nb = torch.empty(B, dtype=torch.int64, pin_memory=True)
    d_bytes = torch.empty(total + 64, dtype=torch.uint8, device=dev)
    d_bytes[:total].copy_(h_bytes[:total], non_blocking=True)
    d_sig = torch.empty(B * S * elem, dtype=torch.uint8, device=dev)
    d_nb = torch.empty(B, dtype=torch.int64, device=dev)
    d_all = torch.empty(world * B * S * elem, dtype=torch.uint8, device=dev) if world > 1 else None
    stream = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()

    def step():
        sk.sketch_device(d_bytes.data_ptr(), offs, B, d_sig.data_ptr(), d_nb.data_ptr(), stream)
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_sig)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, a.warmup)):
        step()
    barrier()

    # ---- timed region: HBM-resident inputs
    sk.enable_timing(True)
    launches0 = sk.launch_count
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        step()
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
    value = world * B * a.steps / (ms / 1e3)

    # sanity, outside the timed region: every signature is filled and encoded lengths are right
    nb = d_nb.cpu().numpy()
    if a.aa:
        assert (nb > 0.5 * a.nprot * 333).all() and (nb < 2 * a.nprot * 333).all(), nb[:4]
    else:
        assert (nb > 0.99 * a.genome_len - 100).all() and (nb <= a.genome_len).all(), nb[:4]
    sig0 = d_sig.cpu().numpy().view(sk.dtype).reshape(B, S)
    assert (sig0 != 0).mean() > 0.999

    # ---- e2e: host pointers through the public C-ABI call (H2D + kernels + D2H every step)
    e2e_steps = max(2, min(a.steps, 4))
    sk.sketch_pointers(h_bytes.data_ptr(), offs, B, h_sig.data_ptr(), h_nb.data_ptr())  # warm staging buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sk.sketch_pointers(h_bytes.data_ptr(), offs, B, h_sig.data_ptr(), h_nb.data_ptr())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(t.item())
    # the same H2D copy alone (pinned -> device): what the e2e number is bounded by on this box
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    d_bytes[:total].copy_(h_bytes[:total], non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_gbs = total / 1e9 / (c0.elapsed_time(c1) / 1e3)
    assert np.array_equal(h_sig.numpy().view(sk.dtype).reshape(B, S), sig0), "e2e and resident paths differ"

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        fasta_per_genome = total / B
        alg_bytes_per_genome = fasta_per_genome + S * elem  # L_fasta + S*sizeof(Sig)
        # The sketch kernels of two genome groups and the FASTA packer of the next group run
        # concurrently on three streams, so a single kernel's own duration is not separable; the
        # span timed here (CUDA events on the launching stream, first group start -> last group
        # end) covers the whole K1 || K2 || K3 pipeline of a call, K2 being ~75 % of its stream time.
        k2_ms, k2_n = ktimes["k2_scan"]
        traffic = measured_traffic("sketch_prob_k21_s18000") if (algo == 0 and not a.aa) else None
        genomes_timed = B * a.steps
        achieved = (alg_bytes_per_genome * genomes_timed / 1e9) / (k2_ms / 1e3) if k2_ms > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(3, a.warmup), "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64" if elem == 8 else ("f32" if algo else "u32"),
            "data": "synthetic",
            "config": {"workload": workload_name(a), "batch_per_gpu": B, "fasta_bytes_per_step_per_gpu": int(total),
                       "l2": "inputs larger than L2 (no flush needed)",
                       "collective": "NCCL all_gather_into_tensor of signatures" if world > 1 else "none"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(total),
                    "d2h_bytes_per_step": int(B * S * elem + B * 8), "steps": e2e_steps,
                    "h2d_copy_alone_gbs": h2d_gbs,
                    "h2d_bound_genomes_per_s": world * B / (total / 1e9 / h2d_gbs)},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {
                "kernel": "prob sketch pipeline: k2_prob_mark/classify/exact (dominant) overlapped with K1 pack and K3"
                          if algo == 0 else "k2_optdens",
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None,
                "traffic": (traffic["dram_bytes_per_genome"] * B if traffic else None),
                "traffic_source": "profiles/r1_traffic.json: ncu dram bytes of all sketch kernels per genome x "
                                  "genomes per step" if traffic else None,
                "peak_source": peak_src, "algorithmic_bytes_per_genome": alg_bytes_per_genome,
                "algorithmic_bytes_per_launch": alg_bytes_per_genome * B,
                "avg_launch_ms": (k2_ms / k2_n) if k2_n else None, "launches_timed": k2_n,
                "note": "L2-atomic / integer-ALU bound, not HBM bound (DESIGN.md); fraction reported as required; "
                        "kernel_ms holds the per-kernel stream times (they overlap across streams)",
            },
            "kernel_ms": {k: v[0] for k, v in ktimes.items()},
            "retries": int(sk.retry_count),
        }
        if world == 1:
            line["cpu_baseline"] = cpu_baseline(a)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(a):
    """the oracle (C port of the reference CPU path) on this box's host cores, bounded sample"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    cores = host_threads()
    sample = max(1, min(a.batch, a.cpu_sample))
    buf, offs = gen_batch_aa_numpy(0, sample, a.nprot) if a.aa else gen_batch_numpy(0, sample, a.genome_len)
    algo = ALGO_ID[a.algo]
    t0 = time.perf_counter()
    O.sketch_buffer(buf, offs, a.kmer, a.sketch, algo, 1 if a.aa else 0, False, 0, nthreads=cores)
    dt = time.perf_counter() - t0
    what = f"{a.nprot}-protein proteomes" if a.aa else f"{a.genome_len} bp genomes"
    return {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample} x {what}, one per thread (oracle/, C restatement)"}


# --------------------------------------------------------------------------- request workload
def tree_signatures(torch, n, S, dev, seed):
    """synthetic u64 signatures with graded distances (a random recursive tree: every point keeps
    a random 50-95 % of the slots of a random earlier point), generated on the device"""
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    base = torch.randint(1, 2**40, (n, S), dtype=torch.int64, device=dev, generator=gen)
    par = (torch.rand(n, device=dev, generator=gen) * torch.arange(n, device=dev)).long().clamp_(min=0)
    keep = 0.5 + 0.45 * torch.rand(n, device=dev, generator=gen)
    par_h, keep_h = par.tolist(), keep.tolist()
    for i in range(1, n):
        m = torch.rand(S, device=dev, generator=gen) < keep_h[i]
        base[i] = torch.where(m, base[par_h[i]], base[i])
    return base


def run_request(a):
    """queries/sec of `request` on one GPU: index built on device by gsb_index_insert_batch_dev
    (timed: genomes/s inserted), then searched with gsb_index_search_batch (host queries in,
    host neighbours out).  Roofline: every distance evaluation streams one candidate signature
    (S * 8 bytes) from HBM; the kernel reports the evaluations it performed."""
    import torch
    import gsearch_b200 as g

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the library has no CPU path")
    dev = torch.device("cuda", 0)
    S, n, nq = a.sketch, a.db, a.queries
    base = tree_signatures(torch, n, S, dev, 1234)
    gen = torch.Generator(device=dev)
    gen.manual_seed(99)
    pick = torch.randint(0, n, (nq,), device=dev, generator=gen)
    fresh = torch.randint(1, 2**40, (nq, S), dtype=torch.int64, device=dev, generator=gen)
    queries = torch.where(torch.rand(nq, S, device=dev, generator=gen) < 0.9, base[pick], fresh)
    h_q = queries.cpu().numpy().view(np.uint64)
    torch.cuda.synchronize()
    idx = g.Hnsw(g.HnswParams(max_nb_conn=a.nbng, ef=a.ef), S, np.uint64)
    t0 = time.perf_counter()
    idx.insert_device(base.data_ptr(), np.arange(n, dtype=np.uint64))
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    clocks = ClockSampler(0)
    for _ in range(max(1, min(a.warmup, 2))):
        out, cnt, neval = idx.search_raw(h_q, a.knbn, a.ef_search)
    steps = max(1, min(a.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        out, cnt, neval = idx.search_raw(h_q, a.knbn, a.ef_search)
    dt = time.perf_counter() - t0
    clk = clocks.stop()
    qps = nq * steps / dt
    peak, peak_src = measured_peak_gbs()
    bytes_step = float(neval.sum()) * S * 8
    achieved = bytes_step * steps / 1e9 / dt
    # recall against brute force on a few queries (size-independent sanity inside the bench)
    d = g.DistHamming().matrix(h_q[:8], base.cpu().numpy().view(np.uint64))
    hits = sum(int((out["distance"][i][:cnt[i]] <= np.sort(d[i])[a.knbn - 1]).sum()) for i in range(8))
    line = {
        "metric": "queries/sec (request)", "value": qps, "unit": "queries/s", "n_gpus": 1, "steps": steps,
        "warmup": max(1, min(a.warmup, 2)), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"configs[2]: {nq} queries vs {n}-signature HNSW built on "
                               f"device (s={S} n={a.nbng} ef={a.ef}), ef_search={a.ef_search}, knbn={a.knbn}",
                   "l2": f"index signatures {n * S * 8 / 1e9:.2f} GB, larger than L2"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": int(nq * S * 8),
                "d2h_bytes_per_step": int(nq * a.knbn * 24 + nq * 12),
                "note": "the timed call takes host queries and returns host neighbours"},
        "gpu_launches": steps, "clocks": clk,
        "roofline": {"kernel": "k7_hnsw_search", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_query": bytes_step / nq, "mean_evaluations_per_query": float(neval.mean())},
        "build": {"genomes_per_s_inserted": n / t_build, "seconds": t_build, "kernel": "k8_hnsw_insert_select"},
        "recall_at_knbn_on_8_queries": hits / (8 * a.knbn),
    }
    # CPU arm: the oracle's search on the SAME graph, one query per host thread (parallel_search)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    cores = host_threads()
    h = O.Hnsw(a.nbng, a.ef, S, np.uint64)
    h.import_graph(base.cpu().numpy().view(np.uint64), idx.export_graph())
    cq = max(1, min(nq, a.cpu_queries))
    t0 = time.perf_counter()
    want, wcnt, _ = h.search(h_q[:cq], a.knbn, a.ef_search, nthreads=cores)
    dtc = time.perf_counter() - t0
    assert want["d_id"].tolist() == out["d_id"][:cq].tolist(), "GPU and CPU answers differ on the same graph"
    line["cpu_baseline"] = {"value": cq / dtc, "unit": "queries/s", "cores": cores, "kind": "port",
                            "sample": f"{cq} queries, one per thread, oracle search on the graph the GPU built"}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    args = parse_args()
    if args.workload == "request":
        if int(os.environ.get("RANK", "0")) == 0:
            run_request(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)
