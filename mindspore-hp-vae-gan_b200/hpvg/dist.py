"""Multi-GPU plumbing for the sampling path: one process per GPU, samples sharded by index, ONE
collective — the all-gather of per-sample FID moments — over NCCL (NVLink 5 / NVSwitch).

The reference has no communication backend at all (single device); this is the B200 addition named by `north_star`.
NCCL is called directly through ctypes on OUR stream with OUR device pointers (no torch in the product path).  The
`ncclUniqueId` is exchanged through a file in a rendezvous directory shared by the ranks of ONE node (torchrun's
MASTER_PORT keys the file name), which is all a single 8-GPU box needs.

A `Communicator` only has to provide `rank`, `world`, `all_gather_rows(local_rows) -> all rows` and `barrier()`; the
CPU tests drive the same sharding / gather logic through a gloo-backed implementation that lives in tests/."""
import ctypes
import glob
import os
import time

import numpy as np

from .runtime import F32, HpvgError, Tensor, _s

NCCL_UNIQUE_ID_BYTES = 128
_NCCL_FLOAT32 = 7   # ncclFloat32


class SingleProcess:
    """world == 1: the gather is the identity."""
    rank, world = 0, 1

    def all_gather_rows(self, rows, stream=None):
        return np.asarray(rows.numpy(stream) if hasattr(rows, "numpy") else rows)

    def barrier(self):
        pass


def _find_nccl():
    cands = []
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            for loc in spec.submodule_search_locations:
                cands += glob.glob(os.path.join(loc, "lib", "libnccl.so*"))
    except Exception:
        pass
    cands += ["libnccl.so.2", "libnccl.so"]
    last = None
    for c in cands:
        try:
            return ctypes.CDLL(c)
        except OSError as e:
            last = e
    raise HpvgError("libnccl not found (%s)" % last)


class _UniqueId(ctypes.Structure):
    _fields_ = [("internal", ctypes.c_byte * NCCL_UNIQUE_ID_BYTES)]


class NcclCommunicator:
    """ncclCommInitRank over a file rendezvous; all_gather_rows = ncclAllGather of equal-sized fp32 row blocks."""

    def __init__(self, rank, world, rendezvous_dir="/tmp", key=None, timeout_s=120.0):
        self.rank, self.world = int(rank), int(world)
        self.nccl = _find_nccl()
        n = self.nccl
        n.ncclGetErrorString.restype = ctypes.c_char_p
        n.ncclGetUniqueId.argtypes = [ctypes.POINTER(_UniqueId)]
        n.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _UniqueId, ctypes.c_int]
        n.ncclAllGather.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p,
                                    ctypes.c_void_p]
        n.ncclCommDestroy.argtypes = [ctypes.c_void_p]
        key = key or os.environ.get("MASTER_PORT", "0")
        # all workers of one torchrun launch share the agent as parent: a per-launch nonce against stale files
        path = os.path.join(rendezvous_dir, "hpvg_nccl_id_%s_%d" % (key, os.getppid()))
        uid = _UniqueId()
        if self.rank == 0:
            self._check(n.ncclGetUniqueId(ctypes.byref(uid)), "ncclGetUniqueId")
            tmp = path + ".tmp.%d" % os.getpid()
            with open(tmp, "wb") as f:
                f.write(bytes(uid.internal))
            os.replace(tmp, path)
        else:
            t0 = time.time()
            while not os.path.exists(path):
                if time.time() - t0 > timeout_s:
                    raise HpvgError("NCCL rendezvous timed out waiting for %s" % path)
                time.sleep(0.01)
            with open(path, "rb") as f:
                raw = f.read()
            ctypes.memmove(uid.internal, raw, NCCL_UNIQUE_ID_BYTES)
        self.comm = ctypes.c_void_p()
        self._check(n.ncclCommInitRank(ctypes.byref(self.comm), self.world, uid, self.rank), "ncclCommInitRank")
        self._path = path

    def _check(self, rc, what):
        if rc != 0:
            raise HpvgError("%s failed: %s" % (what, self.nccl.ncclGetErrorString(rc).decode()))

    def all_gather_rows(self, rows, stream=None):
        """rows: device fp32 Tensor (n_local, k), the SAME n_local on every rank.  Returns host (world*n_local, k)."""
        n_local, k = rows.shape
        out = Tensor((self.world * n_local, k), F32)
        self._check(self.nccl.ncclAllGather(rows.ptr, out.ptr, n_local * k, _NCCL_FLOAT32, self.comm, _s(stream)),
                    "ncclAllGather")
        return out.numpy(stream)

    def barrier(self):
        t = Tensor((1, 1), F32).zero_()
        self.all_gather_rows(t)

    def close(self):
        if self.comm:
            self.nccl.ncclCommDestroy(self.comm)
            self.comm = None
        if self.rank == 0:
            try:
                os.remove(self._path)
            except OSError:
                pass


def from_env():
    """Communicator for the current launch: torchrun-style RANK / WORLD_SIZE env, else single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return SingleProcess()
    return NcclCommunicator(int(os.environ["RANK"]), world)


# ---------------------------------------------------------------------------------------------------------------------
# sharding helpers (pure host logic; exercised on CPU with world_size 2 in tests/test_cpu_dist.py)
# ---------------------------------------------------------------------------------------------------------------------
def shard_rows(num_samples, batch, rank, world):
    """Global sample indices owned by `rank` (chunks of `batch` dealt round-robin, see sampling.local_chunks), padded
    with -1 so that every rank holds the same number of rows (ncclAllGather needs equal counts)."""
    from .sampling import local_chunks
    mine = [i for c in local_chunks(num_samples, batch, rank, world) for i in c]
    n_chunks = (num_samples + batch - 1) // batch
    max_chunks = (n_chunks + world - 1) // world
    pad = max_chunks * batch - len(mine)
    return mine + [-1] * pad


def unshard_rows(all_rows, num_samples, batch, world):
    """Inverse of the deal: (world * rows_per_rank, k) gathered rows -> (num_samples, k) in global sample order."""
    all_rows = np.asarray(all_rows)
    per = all_rows.shape[0] // world
    out = np.zeros((num_samples,) + all_rows.shape[1:], all_rows.dtype)
    seen = np.zeros(num_samples, bool)
    for r in range(world):
        idx = shard_rows(num_samples, batch, r, world)
        for j, i in enumerate(idx):
            if i >= 0:
                out[i] = all_rows[r * per + j]
                seen[i] = True
    if not seen.all():
        raise HpvgError("unshard_rows: %d samples missing" % int((~seen).sum()))
    return out
