"""Device memory, streams and events over libhpvg's runtime shim (numpy on the host side, no torch)."""
import ctypes

import numpy as np

from ._lib import HpvgError, check, lib

_initialised = {"device": None}

BF16 = "bfloat16"
F32 = "float32"
F64 = "float64"
I32 = "int32"
U64 = "uint64"
U8 = "uint8"
_ITEMSIZE = {BF16: 2, F32: 4, F64: 8, I32: 4, U64: 8, U8: 1}
_NP = {F32: np.float32, F64: np.float64, I32: np.int32, BF16: np.uint16, U64: np.uint64, U8: np.uint8}


def init(device=0):
    """Select the GPU and create library scratch.  Raises when no sm_100a device is visible (no CPU fallback)."""
    if lib.hpvg_device_count() <= 0:
        raise HpvgError("no CUDA device visible: the hpvg product path requires a B200 (there is no CPU fallback)")
    check(lib.hpvg_init(int(device)), "hpvg_init")
    _initialised["device"] = int(device)


def is_initialised():
    return _initialised["device"] is not None


def sm_count():
    return lib.hpvg_sm_count()


class Stream:
    def __init__(self):
        h = ctypes.c_void_p()
        check(lib.hpvg_stream_create(ctypes.byref(h)), "stream_create")
        self.handle = h

    def sync(self):
        check(lib.hpvg_stream_sync(self.handle), "stream_sync")

    def wait_event(self, event):
        """Work enqueued on this stream after the call waits for `event` (device-side, the host does not block)."""
        check(lib.hpvg_stream_wait_event(self.handle, event.handle), "stream_wait_event")


class Event:
    def __init__(self):
        h = ctypes.c_void_p()
        check(lib.hpvg_event_create(ctypes.byref(h)), "event_create")
        self.handle = h

    def __del__(self):
        try:
            lib.hpvg_event_destroy(self.handle)
        except Exception:
            pass

    def record(self, stream=None):
        check(lib.hpvg_event_record(self.handle, _s(stream)), "event_record")

    def sync(self):
        check(lib.hpvg_event_sync(self.handle), "event_sync")

    def elapsed_ms(self, end):
        ms = ctypes.c_float()
        check(lib.hpvg_event_elapsed_ms(self.handle, end.handle, ctypes.byref(ms)), "event_elapsed")
        return ms.value


def _s(stream):
    return stream.handle if stream is not None else None


def device_sync():
    check(lib.hpvg_device_sync(), "device_sync")


# Caching allocator: cudaMalloc/cudaFree are slow and cudaFree synchronises the device, so freed blocks are kept in
# size-keyed free lists and reused.  Safe because every consumer enqueues on ONE stream per process (in-order), so a
# block is never rewritten before earlier kernels that read it have run.
_POOL = {}
_POOL_STATS = {"malloc": 0, "reuse": 0}


def _pool_alloc(nbytes):
    size = max(512, (int(nbytes) + 511) // 512 * 512)
    lst = _POOL.get(size)
    if lst:
        _POOL_STATS["reuse"] += 1
        ptr = lst.pop()
    else:
        h = ctypes.c_void_p()
        check(lib.hpvg_malloc(ctypes.byref(h), size), "malloc")
        _POOL_STATS["malloc"] += 1
        ptr = h.value
    if _CAPTURE_HOLD[0] is not None:
        # allocated while a graph is being captured: the address is baked into captured kernel arguments, so the block
        # belongs to that graph until it is destroyed — whenever its Python owner lets go of it
        _GRAPH_OWNED[ptr] = _CAPTURE_HOLD[0]
    return ptr, size


_CAPTURE_HOLD = [None]   # while a CUDA graph is being captured: blocks freed meanwhile stay reserved for the graph
_GRAPH_OWNED = {}        # ptr -> held-list of the live graph whose capture allocated it


def _pool_free(ptr, size):
    held = _GRAPH_OWNED.get(ptr)
    if held is not None:
        held.append((ptr, size))         # stays reserved for its graph (returned to the pool by Graph.destroy)
    elif _CAPTURE_HOLD[0] is not None:
        _CAPTURE_HOLD[0].append((ptr, size))
    else:
        _POOL.setdefault(size, []).append(ptr)


def _sync_unless_capturing():
    """cudaDeviceSynchronize is illegal while a stream is being captured into a graph."""
    if _CAPTURE_HOLD[0] is None:
        device_sync()


def empty_cache():
    device_sync()
    for lst in _POOL.values():
        for ptr in lst:
            lib.hpvg_free(ctypes.c_void_p(ptr))
    _POOL.clear()


class Tensor:
    """A shaped view of caller-owned device memory.  dtype in {float32, bfloat16, float64, int32}."""

    __slots__ = ("ptr", "shape", "dtype", "_owner", "nbytes", "_cap")

    def __init__(self, shape, dtype=F32, ptr=None, owner=None):
        self.shape = tuple(int(v) for v in shape)
        self.dtype = dtype
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * _ITEMSIZE[dtype] if len(self.shape) else _ITEMSIZE[dtype]
        if ptr is None:
            if not is_initialised():
                raise HpvgError("hpvg.init() must be called before allocating device memory")
            self.ptr, self._cap = _pool_alloc(self.nbytes)
            self._owner = True
        else:
            self.ptr = int(ptr)
            self._cap = 0
            self._owner = owner   # keeps the parent allocation alive

    def __del__(self):
        try:
            if self._owner is True and self.ptr:
                _pool_free(self.ptr, self._cap)
        except Exception:
            pass

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64)) if len(self.shape) else 1

    def view(self, shape, dtype=None, byte_offset=0):
        return Tensor(shape, dtype or self.dtype, ptr=self.ptr + byte_offset, owner=self)

    # Stream contract of the data-movement methods.  hpvg.Stream objects are cudaStreamNonBlocking: they do NOT order
    # against the legacy NULL stream.  `stream=None` therefore means "synchronous with respect to the whole device":
    # the call first waits for everything queued on every stream (so it sees the results of kernels a caller enqueued on
    # its own stream — e.g. checkpoint.state_dict() after driver.train_scale(stream=st)) and returns only when its own
    # copy has completed (so kernels enqueued afterwards on any stream see the data).  With an explicit stream the call
    # is ordered on that stream only.
    def zero_(self, stream=None):
        if stream is None:
            _sync_unless_capturing()
        check(lib.hpvg_memset(self.ptr, 0, self.nbytes, _s(stream)), "memset")
        if stream is None:
            _sync_unless_capturing()
        return self

    def copy_from_host(self, arr, stream=None):
        a = np.ascontiguousarray(arr, dtype=_NP[self.dtype])
        if a.nbytes != self.nbytes:
            raise HpvgError("copy_from_host: size mismatch %d vs %d" % (a.nbytes, self.nbytes))
        if stream is None:
            _sync_unless_capturing()
        check(lib.hpvg_h2d(self.ptr, a.ctypes.data, self.nbytes, _s(stream)), "h2d")
        if stream is not None:
            stream.sync()   # pageable source must stay alive until the copy is done
        else:
            _sync_unless_capturing()
        return self

    def numpy(self, stream=None):
        out = np.empty(self.shape, dtype=_NP[self.dtype])
        if stream is None:
            device_sync()   # producers may sit on any (non-blocking) stream
        check(lib.hpvg_d2h(out.ctypes.data, self.ptr, self.nbytes, _s(stream)), "d2h")
        if stream is not None:
            stream.sync()
        else:
            device_sync()
        return out

    def copy_(self, other, stream=None):
        if other.nbytes != self.nbytes:
            raise HpvgError("copy_: size mismatch")
        if stream is None:
            _sync_unless_capturing()
        check(lib.hpvg_d2d(self.ptr, other.ptr, self.nbytes, _s(stream)), "d2d")
        if stream is None:
            _sync_unless_capturing()
        return self


class Graph:
    """CUDA graph of everything enqueued on `stream` inside the `with` block.  Device blocks that are released to the
    allocator during capture stay reserved for the graph (their addresses are baked into the captured kernel
    arguments) until `destroy()`.  After capture, drive the captured objects ONLY through `launch()`."""

    def __init__(self, stream):
        self.stream, self.exec, self._held = stream, None, []

    def __enter__(self):
        if _CAPTURE_HOLD[0] is not None:
            raise HpvgError("nested graph capture")
        _CAPTURE_HOLD[0] = self._held
        check(lib.hpvg_graph_begin(self.stream.handle), "graph_begin")
        return self

    def __exit__(self, et, ev, tb):
        _CAPTURE_HOLD[0] = None
        h = ctypes.c_void_p()
        rc = lib.hpvg_graph_end(self.stream.handle, ctypes.byref(h))
        if et is None:
            check(rc, "graph_end")
            self.exec = h
        return False

    def launch(self):
        check(lib.hpvg_graph_launch(self.exec, self.stream.handle), "graph_launch")

    def destroy(self):
        if self.exec is not None:
            self.stream.sync()
            lib.hpvg_graph_destroy(self.exec)
            self.exec = None
        held = self._held
        for ptr in [q for q, h in _GRAPH_OWNED.items() if h is held]:
            del _GRAPH_OWNED[ptr]        # blocks still alive simply return to the pool when their owner frees them
        for ptr, size in held:
            _POOL.setdefault(size, []).append(ptr)
        self._held = []


def from_numpy(arr, dtype=None, stream=None):
    a = np.asarray(arr)
    if dtype is None:
        dtype = {np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.int32): I32,
                 np.dtype(np.uint64): U64, np.dtype(np.uint8): U8}.get(a.dtype, F32)
    t = Tensor(a.shape, dtype)
    t.copy_from_host(a, stream)
    return t


def bf16_bits_to_f32(u16):
    return (np.asarray(u16, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


class PinnedBuffer:
    """Page-locked host staging buffer (for timed H2D/D2H in bench.py)."""

    def __init__(self, nbytes):
        h = ctypes.c_void_p()
        check(lib.hpvg_host_alloc(ctypes.byref(h), int(nbytes)), "host_alloc")
        self.ptr = h.value
        self.nbytes = int(nbytes)

    def as_array(self, shape, dtype=np.float32):
        n = int(np.prod(shape))
        buf = (ctypes.c_byte * (n * np.dtype(dtype).itemsize)).from_address(self.ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def __del__(self):
        try:
            lib.hpvg_host_free(ctypes.c_void_p(self.ptr))
        except Exception:
            pass
