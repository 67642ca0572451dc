"""Per-device contexts and the glue between torch tensors and the C ABI.

torch is plumbing only: it owns device memory (caching allocator), the current
stream and torch.distributed.  All arithmetic happens in libshortseq_b200.so.
"""
import ctypes as C
import threading

import numpy as np
import torch

from . import _lib
from ._lib import LibraryError, Report

_CTX = {}
_LOCK = threading.Lock()

# messages of the reference (short_seq.pyx:74, short_seq_64.pyx:78-80,105)
MSG_TOO_LONG = "Sequences longer than 1024 bases are not supported."


class ShortSeqClassError(ValueError):
    """A read's length does not belong to the container class the batch call was made for."""


class Context:
    """One GPU: an ssq_ctx bound to torch's current stream on that device."""

    def __init__(self, device_index):
        if not torch.cuda.is_available():
            raise LibraryError("no CUDA device available: shortseq_b200 runs on the GPU only (no CPU fallback)")
        self.device = torch.device("cuda", device_index)
        self.index = device_index
        h = C.c_void_p()
        _lib.check(_lib.lib().ssq_ctx_create(device_index, C.byref(h)))
        self.handle = h

    def bind(self):
        """Enqueue on torch's current stream of this device so kernels order with torch ops."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().ssq_ctx_set_stream(self.handle, C.c_void_p(stream)))
        return self.handle

    def sync(self):
        rep = Report()
        _lib.check(_lib.lib().ssq_ctx_sync(self.handle, C.byref(rep)))
        return rep

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)


def context(device=None):
    """The (cached) context of a device; `device` is None (current), an int or a torch.device."""
    if device is None:
        if not torch.cuda.is_available():
            raise LibraryError("no CUDA device available: shortseq_b200 runs on the GPU only (no CPU fallback)")
        idx = torch.cuda.current_device()
    elif isinstance(device, int):
        idx = device
    else:
        device = torch.device(device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
    with _LOCK:
        ctx = _CTX.get(idx)
        if ctx is None:
            _lib.lib()  # loud failure if the library is missing
            ctx = _CTX[idx] = Context(idx)
    return ctx


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def to_device(ctx, a, dtype=None):
    """numpy array / tensor -> contiguous tensor on the context's device."""
    if isinstance(a, torch.Tensor):
        t = a
    else:
        a = np.ascontiguousarray(a)
        if a.dtype == np.uint64:
            a = a.view(np.int64)
        elif a.dtype == np.uint16:
            a = a.view(np.int16)
        t = torch.from_numpy(a)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(ctx.device, non_blocking=False).contiguous()


def words_to_numpy(t):
    """int64 tensor holding uint64 bit patterns -> numpy uint64 array (host)."""
    return t.detach().cpu().numpy().view(np.uint64)


_FASTBOX = False


def fastbox():
    """The CPython helper module (hostext/_fastbox.c: gather / boxing / dict-fill loops in C), or None when it has not
    been built -- the same loops then run in Python.  Host-side glue only: no sequence arithmetic lives there."""
    global _FASTBOX
    if _FASTBOX is False:
        try:
            from . import _fastbox as m
            _FASTBOX = m
        except ImportError:
            _FASTBOX = None
    return _FASTBOX


def gather_reads(reads):
    """list of bytes -> (uint8 buffer, int64 offsets[n+1]) on the host.

    Mirrors the element check of the reference's `<bytes> PyList_GET_ITEM` cast
    (counter.pyx:27): a non-bytes element is a TypeError.
    """
    fb = fastbox()
    if fb is not None and type(reads) is list:
        buf, off = fb.gather(reads)                     # one C pass for the sizes, one memcpy pass
        return np.frombuffer(buf, dtype=np.uint8), np.frombuffer(off, dtype=np.int64)
    n = len(reads)
    kinds = set(map(type, reads))                       # C-speed pass; the common case is {bytes}
    if kinds - {bytes}:
        bad = next(r for r in reads if type(r) is not bytes)
        raise TypeError(f"expected bytes, {type(bad).__name__} found")
    lens = np.fromiter(map(len, reads), dtype=np.int64, count=n)
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    buf = np.frombuffer(b"".join(reads), dtype=np.uint8)
    return buf, offsets


def bad_base_message(read: bytes) -> str:
    """The message the reference raises for this read (SURVEY trap T2).

    ShortSeq64 reads and block tails are scanned from the end and report the last bad
    character (short_seq_64.pyx:101-105, util.pyx:133-137); full 32-nt blocks are checked
    first, block by block, 8-byte chunks j=3..0, and report the whole chunk (util.pyx:113-115).
    Validation here is exact {A,C,G,T}; the reference's bloom filter additionally lets 16
    alias byte values through (documented divergence, SURVEY trap T1).
    """
    ok = b"ACGT"
    n = len(read)
    if n > 32:
        for blk in range(n // 32):
            for j in (3, 2, 1, 0):
                chunk = read[32 * blk + 8 * j: 32 * blk + 8 * j + 8]
                if any(c not in ok for c in chunk):
                    return f"Unsupported base character: {chunk.decode('ascii', errors='replace')}"
        tail = read[32 * (n // 32):]
    else:
        tail = read
    for c in reversed(tail):
        if c not in ok:
            return f"Unsupported base character: {bytes([c]).decode('ascii', errors='replace')}"
    return "Unsupported base character"
