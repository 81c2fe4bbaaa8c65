"""hnsw_rs `hnswio`-style dump / reload of an index, behind GSB_DUMP_HNSWIO=1.

The reference writes its database with `hnsw.file_dump(dir, "hnswdump")` (src/utils/dumpload.rs:31) and
reloads it with `HnswIo::load_hnsw` after sniffing `Description.t_name` (src/utils/reloadhnsw.rs:13-37,
src/bin/gsearch.rs:807-851).  The byte layout of those two files lives in the hnsw_rs crate, which is
NOT in the reference tree; what is written here follows the layout SURVEY A.11 recalls:

  <base>.hnsw.graph   u32 magic 0x000a677f | u8 dump mode (1 = with data) | u8 max_nb_connection |
                      u8 nb_layer | u64 ef | u64 nb_point | u64 data dimension | (u64 len, utf8) distname |
                      (u64 len, utf8) t_name in {"u16","u32","u64","f32"} | then layer by layer:
                      u64 points in the layer, per point u32 magic 0x000a678f, u64 origin_id, p_id = (u8 layer,
                      i32 rank), and for each of the nb_layer layers a neighbour count followed by
                      (u64 origin_id, u8 layer, i32 rank, f32 distance) | entry point (u64 origin_id, u8, i32)
  <base>.hnsw.data    u32 magic 0xa67f0000 | u64 dimension | per point, same order: u32 magic, u64 origin_id,
                      u64 byte length, raw little-endian elements

BYTE COMPATIBILITY WITH STOCK hnsw_rs IS UNVERIFIED (no upstream source, no Rust toolchain here).  One
deliberate deviation: A.11 recalls a u8 neighbour count, which cannot hold the 256 links a layer-0 list
has at the reference's own default `-n 128`; the count is written as u16.  The native layout
(gsb_index_dump) stays the default; this module exists so that the conversion is
one function away once the upstream layout can be checked.  Everything here is host-side file I/O over
the graph image of gsb_index_export_graph / gsb_index_load_graph."""
import os
import struct

import numpy as np

MAGIC_DESCR = 0x000A677F
MAGIC_POINT = 0x000A678F
MAGIC_DATA = 0xA67F0000
NB_LAYER = 16
T_NAMES = {np.dtype(np.uint16): "u16", np.dtype(np.uint32): "u32", np.dtype(np.uint64): "u64",
           np.dtype(np.float32): "f32"}
DISTNAME = "DistHamming"


def is_hnswio(directory, basename="hnswdump"):
    try:
        with open(os.path.join(directory, basename + ".hnsw.graph"), "rb") as f:
            return struct.unpack("<I", f.read(4))[0] == MAGIC_DESCR
    except (OSError, struct.error):
        return False


def dump(directory, basename, image, sigs, max_nb_connection, ef):
    """image: the dict of Hnsw.export_graph() (levels, ranks, ids, nbr_offsets, nbr_index, nbr_dist,
    entry_point); sigs: n x S array"""
    levels, ranks, ids = image["levels"], image["ranks"], image["ids"]
    off, nidx, ndist = image["nbr_offsets"], image["nbr_index"], image["nbr_dist"]
    n = len(ids)
    sigs = np.ascontiguousarray(sigs)
    S = sigs.shape[1] if n else 0
    first_list = np.zeros(n + 1, dtype=np.int64)           # list (p, l) is number first_list[p] + l
    np.cumsum(levels.astype(np.int64) + 1, out=first_list[1:])
    nbr_rec = np.dtype([("id", "<u8"), ("layer", "u1"), ("rank", "<i4"), ("dist", "<f4")])
    with open(os.path.join(directory, basename + ".hnsw.graph"), "wb") as g, \
            open(os.path.join(directory, basename + ".hnsw.data"), "wb") as d:
        g.write(struct.pack("<IBBBQQQ", MAGIC_DESCR, 1, int(max_nb_connection), NB_LAYER, int(ef), n, S))
        for name in (DISTNAME, T_NAMES[sigs.dtype]):
            g.write(struct.pack("<Q", len(name)) + name.encode())
        d.write(struct.pack("<IQ", MAGIC_DATA, S))
        for layer in range(NB_LAYER):
            members = np.nonzero(levels == layer)[0]
            members = members[np.argsort(ranks[members], kind="stable")]
            g.write(struct.pack("<Q", len(members)))
            for p in members:
                g.write(struct.pack("<IQBi", MAGIC_POINT, int(ids[p]), layer, int(ranks[p])))
                for l in range(NB_LAYER):
                    if l > layer:
                        g.write(struct.pack("<H", 0))
                        continue
                    a, b = int(off[first_list[p] + l]), int(off[first_list[p] + l + 1])
                    nb = nidx[a:b]
                    rec = np.empty(b - a, dtype=nbr_rec)
                    rec["id"], rec["layer"], rec["rank"], rec["dist"] = ids[nb], levels[nb], ranks[nb], ndist[a:b]
                    g.write(struct.pack("<H", b - a) + rec.tobytes())
                raw = sigs[p].tobytes()
                d.write(struct.pack("<IQQ", MAGIC_DATA, int(ids[p]), len(raw)) + raw)
        e = int(image["entry_point"]) if n else 0
        g.write(struct.pack("<QBi", int(ids[e]) if n else 0, int(levels[e]) if n else 0, int(ranks[e]) if n else 0))


def load(directory, basename="hnswdump"):
    """-> dict(max_nb_connection, ef, dtype, sigs, ids, levels, ranks, nbr_offsets, nbr_index, nbr_dist,
    entry_point); points are numbered in file order (layer by layer, by rank)"""
    # (memory-mapped: the data file is the whole signature matrix)
    g = np.memmap(os.path.join(directory, basename + ".hnsw.graph"), dtype=np.uint8, mode="r")
    d = np.memmap(os.path.join(directory, basename + ".hnsw.data"), dtype=np.uint8, mode="r")
    magic, mode, M, nb_layer, ef, n, S = struct.unpack_from("<IBBBQQQ", g, 0)
    if magic != MAGIC_DESCR or nb_layer != NB_LAYER:
        raise ValueError("not an hnswio graph file")
    pos = struct.calcsize("<IBBBQQQ")
    names = []
    for _ in range(2):
        (ln,) = struct.unpack_from("<Q", g, pos)
        names.append(bytes(g[pos + 8:pos + 8 + ln]).decode())
        pos += 8 + ln
    dtype = {v: k for k, v in T_NAMES.items()}[names[1]]
    dmagic, dS = struct.unpack_from("<IQ", d, 0)
    if dmagic != MAGIC_DATA or dS != S:
        raise ValueError("not an hnswio data file")
    dpos = 12
    ids = np.zeros(n, dtype=np.uint64)
    levels = np.zeros(n, dtype=np.uint8)
    ranks = np.zeros(n, dtype=np.uint32)
    sigs = np.zeros((n, S), dtype=dtype)
    lists = []                                   # per point: list of (array of (layer, rank), array of dist) per layer
    index_of = {}
    nbr_rec = np.dtype([("id", "<u8"), ("layer", "u1"), ("rank", "<i4"), ("dist", "<f4")])
    p = 0
    for layer in range(NB_LAYER):
        (cnt,) = struct.unpack_from("<Q", g, pos)
        pos += 8
        for _ in range(cnt):
            pm, oid, lay, rk = struct.unpack_from("<IQBi", g, pos)
            pos += struct.calcsize("<IQBi")
            if pm != MAGIC_POINT or lay != layer:
                raise ValueError("corrupt point record")
            ids[p], levels[p], ranks[p] = oid, lay, rk
            index_of[(lay, rk)] = p
            mine = []
            for l in range(NB_LAYER):
                (k,) = struct.unpack_from("<H", g, pos)
                pos += 2
                rec = np.frombuffer(g, dtype=nbr_rec, count=k, offset=pos)
                pos += k * nbr_rec.itemsize
                if l <= layer:
                    mine.append(rec)
            lists.append(mine)
            dm, did, nbytes = struct.unpack_from("<IQQ", d, dpos)
            dpos += 20
            if dm != MAGIC_DATA or did != oid or nbytes != S * np.dtype(dtype).itemsize:
                raise ValueError("corrupt data record")
            sigs[p] = np.frombuffer(d, dtype=dtype, count=S, offset=dpos)
            dpos += nbytes
            p += 1
    if p != n:
        raise ValueError("point count mismatch")
    _, elay, erk = struct.unpack_from("<QBi", g, pos)
    total = sum(len(m) for m in lists)
    off = np.zeros(total + 1, dtype=np.uint64)
    nidx, ndist = [], []
    li = 0
    for mine in lists:
        for rec in mine:
            nidx.append(np.fromiter((index_of[(int(a), int(b))] for a, b in zip(rec["layer"], rec["rank"])),
                                    dtype=np.uint32, count=len(rec)))
            ndist.append(rec["dist"].astype(np.float32))
            off[li + 1] = off[li] + len(rec)
            li += 1
    return dict(max_nb_connection=M, ef=ef, dtype=dtype, sigs=sigs, ids=ids, levels=levels, ranks=ranks,
                nbr_offsets=off, nbr_index=np.concatenate(nidx) if nidx else np.zeros(0, np.uint32),
                nbr_dist=np.concatenate(ndist) if ndist else np.zeros(0, np.float32),
                entry_point=index_of[(elay, erk)] if n else 0)
