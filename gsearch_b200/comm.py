"""Multi-GPU plumbing of the path behind the C ABI: one process per GPU, NCCL over NVLink.

The reference is a single process (src/dna/dnasketch.rs:421-435); its units of parallel work are one
genome per sketch task and one query per search task, which is how the work shards here:

  * `tohnsw --gpus N`: genome i belongs to rank i mod N; each rank sketches its genomes straight into
    its slice of a device matrix, ONE all-gather (`Comm.all_gather_rows`) puts every signature on every
    GPU in file order, then `Hnsw.insert_sharded` builds the graph with the distance evaluations of
    every insertion wave spread over the GPUs (all replicas end identical);
  * `request --gpus N`: replicated index, query j belongs to rank j mod N.

No torch here: device buffers, the communicator and the collectives are the library's (gsb_comm_*,
gsb_device_*).  The NCCL unique id travels through whatever the host has -- a multiprocessing pipe in
cli.py."""
import ctypes as C

import numpy as np

from . import _lib


class DeviceBuffer:
    """cudaMalloc'ed bytes on one GPU (gsb_device_malloc)"""

    def __init__(self, nbytes, device=0):
        self.device, self.nbytes = device, int(nbytes)
        p = C.c_void_p()
        _lib.check(_lib.lib().gsb_device_malloc(device, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def upload(self, array, offset=0):
        a = np.ascontiguousarray(array)
        assert offset + a.nbytes <= self.nbytes
        _lib.check(_lib.lib().gsb_memcpy_h2d(self.device, C.c_void_p(self.ptr + offset), C.c_void_p(a.ctypes.data),
                                             a.nbytes))

    def download(self, dtype, count, offset=0):
        out = np.empty(count, dtype=dtype)
        assert offset + out.nbytes <= self.nbytes
        _lib.check(_lib.lib().gsb_memcpy_d2h(self.device, C.c_void_p(out.ctypes.data), C.c_void_p(self.ptr + offset),
                                             out.nbytes))
        return out

    def free(self):
        if getattr(self, "ptr", None):
            _lib.lib().gsb_device_free(self.device, C.c_void_p(self.ptr))
            self.ptr = None

    __del__ = free


class PinnedBuffer:
    """page-locked host bytes (gsb_host_alloc_pinned): H2D copies from it overlap kernels"""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        _lib.check(_lib.lib().gsb_host_alloc_pinned(self.nbytes, C.byref(p)))
        self.ptr = p.value
        self.array = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(max(self.nbytes, 1),))

    def free(self):
        if getattr(self, "ptr", None):
            self.array = None
            _lib.lib().gsb_host_free_pinned(C.c_void_p(self.ptr))
            self.ptr = None

    __del__ = free


def unique_id():
    """rank 0 draws it; every rank passes the same bytes to Comm()"""
    buf = (C.c_uint8 * 128)()
    _lib.check(_lib.lib().gsb_comm_unique_id(buf))
    return bytes(buf)


def shard_indices(n, rank, world):
    """global indices of the units (genomes or queries) owned by `rank`"""
    return list(range(rank, n, world))


def shard_rows(n, world):
    """rows every rank contributes to the gather (shards are padded to the largest one)"""
    return (n + world - 1) // world


class Comm:
    def __init__(self, uid: bytes, world: int, rank: int, device: int):
        assert len(uid) == 128
        h = C.c_void_p()
        ub = (C.c_uint8 * 128).from_buffer_copy(uid)
        _lib.check(_lib.lib().gsb_comm_create(ub, world, rank, device, C.byref(h)))
        self._h, self.world, self.rank, self.device = h, world, rank, device

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().gsb_comm_destroy(self._h)
            self._h = None

    __del__ = close

    def all_gather_rows(self, d_local_ptr, rows_per_rank, row_bytes, n_total, d_tmp_ptr, d_out_ptr):
        """unit i (on rank i mod world, local row i // world) -> row i of d_out on every rank"""
        _lib.check(_lib.lib().gsb_comm_all_gather_rows(self._h, C.c_void_p(d_local_ptr), rows_per_rank, row_bytes,
                                                       n_total, C.c_void_p(d_tmp_ptr), C.c_void_p(d_out_ptr),
                                                       C.c_void_p(0)))

    def all_gather(self, d_send_ptr, d_recv_ptr, bytes_per_rank):
        _lib.check(_lib.lib().gsb_comm_all_gather(self._h, C.c_void_p(d_send_ptr), C.c_void_p(d_recv_ptr),
                                                  bytes_per_rank, C.c_void_p(0)))

    def broadcast(self, d_ptr, nbytes, root=0):
        _lib.check(_lib.lib().gsb_comm_broadcast(self._h, C.c_void_p(d_ptr), nbytes, root, C.c_void_p(0)))
