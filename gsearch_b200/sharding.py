"""Multi-GPU layout of the path (one process per GPU, torch.distributed for the plumbing).

The reference has no distributed code: its units of parallel work are one genome per sketch
task (src/dna/dnasketch.rs:325-364) and one query per search task (parallel_search,
src/dna/dnarequest.rs:353).  Both shard without any data-path exchange:

  * sketching: genome i belongs to rank i mod G; after sketching, ONE all-gather puts every
    finished signature on every rank, in global genome order, because `tohnsw` inserts them all
    into one graph (src/dna/dnasketch.rs:421-435);
  * request: the graph and signatures are replicated, query j belongs to rank j mod G, and the
    (tiny) answers are gathered.
HNSW insertion itself does not shard (one mutable graph): single builder, replicas only.

The functions below are backend-agnostic (NCCL on the GPU box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_indices(n, rank, world):
    """global indices of the units (genomes or queries) owned by `rank`"""
    return list(range(rank, n, world))


def shard_rows(n, world):
    """rows every rank contributes to the gather (shards are padded to the largest one)"""
    return (n + world - 1) // world


def gather_rows(local, n_total, rank=None, world=None, group=None, out=None):
    """All-gather row-sharded results into global order.

    local   [n_local, ...] rows of the units `shard_indices(n_total, rank, world)`, in that order
    returns [n_total, ...] on every rank, row i = unit i
    `out` may be a preallocated [world * shard_rows, ...] staging tensor (bench: no allocation
    inside the timed region)."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return local[:n_total]
    per = shard_rows(n_total, world)
    tail = local.shape[1:]
    if local.shape[0] != per:  # ragged last round: pad with zero rows
        pad = torch.zeros((per,) + tuple(tail), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
        local = pad
    stage = out if out is not None else torch.empty((world * per,) + tuple(tail), dtype=local.dtype,
                                                     device=local.device)
    dist.all_gather_into_tensor(stage, local.contiguous(), group=group)
    # stage[r * per + t] is unit t * world + r: interleave the shards back into global order
    full = stage.view((world, per) + tuple(tail)).transpose(0, 1).reshape((world * per,) + tuple(tail))
    return full[:n_total]


def sketch_sharded(sketch_fn, files, rank, world, device="cpu", group=None, row_bytes=None, sig_dtype=None):
    """`tohnsw` sketch phase over G ranks.  sketch_fn(list of files) -> (sig [n, S] numpy, nb_bases [n]);
    returns the full (signatures, nb_bases) in file order on every rank.  A rank whose shard is
    empty (fewer files than ranks) still takes part in the gather: it needs `row_bytes` (S *
    sizeof(Sig)) and `sig_dtype`, which it cannot learn from an empty result."""
    import numpy as np

    mine = shard_indices(len(files), rank, world)
    if mine:
        sig, nb = sketch_fn([files[i] for i in mine])
        sig = np.ascontiguousarray(sig)
        row_bytes = sig.dtype.itemsize * sig.shape[1]
        sig_dtype = sig.dtype
    else:
        if row_bytes is None or sig_dtype is None:
            raise ValueError("sketch_sharded: this rank owns no file; pass row_bytes and sig_dtype")
        sig, nb = np.zeros((0, row_bytes // np.dtype(sig_dtype).itemsize), dtype=sig_dtype), np.zeros(0, np.uint64)
    sig_t = torch.from_numpy(sig.view(np.uint8).reshape(len(mine), row_bytes)).to(device)
    nb_t = torch.from_numpy(np.asarray(nb).astype(np.int64)).to(device)
    all_sig = gather_rows(sig_t, len(files), rank, world, group)
    all_nb = gather_rows(nb_t, len(files), rank, world, group)
    sig_all = all_sig.cpu().numpy().view(sig_dtype).reshape(len(files), -1)
    return sig_all, all_nb.cpu().numpy().astype(np.uint64)


def search_sharded(search_fn, queries, knbn, rank, world, device="cpu", group=None):
    """`request` over G ranks with a replicated index.  search_fn(queries) -> (structured
    neighbour array [nq, knbn], counts [nq]); returns both for all queries on every rank."""
    import numpy as np

    from .index import Neighbour

    mine = shard_indices(len(queries), rank, world)
    if mine:
        out, counts = search_fn(queries[mine])
    else:  # fewer queries than ranks: contribute an empty shard, do not call the index
        out, counts = np.zeros((0, knbn), dtype=Neighbour), np.zeros(0, np.uint32)
    item = out.dtype.itemsize
    out_t = torch.from_numpy(np.ascontiguousarray(out).view(np.uint8).reshape(len(mine), knbn * item)).to(device)
    cnt_t = torch.from_numpy(np.asarray(counts).astype(np.int64)).to(device)
    all_out = gather_rows(out_t, len(queries), rank, world, group)
    all_cnt = gather_rows(cnt_t, len(queries), rank, world, group)
    res = all_out.cpu().numpy().view(out.dtype).reshape(len(queries), knbn)
    return res, all_cnt.cpu().numpy().astype(np.uint32)
