"""The reference's per-object API: pack / from_str / from_bytes and ShortSeq64 / ShortSeq192 / ShortSeqVar.

Objects are immutable holders of (packed words, length) with the reference's protocol
(reference short_seq_64.pyx:33-90, short_seq_192.pyx:27-97, short_seq_var.pyx:15-93,
short_seq.pyx:13-238).  Packing, decoding and the Hamming operator run on the GPU -- as
batches of one through the library's single-object entry points (ssq_pack_one / ssq_decode_one /
ssq_hamming_one: one small kernel launch each, operands and results in a page of mapped pinned
memory, no tensors); equality, hashing, length and slicing are integer bookkeeping on the held
words and stay on the host, as SURVEY section 8 row F14 assigns them.
"""
import ctypes as _C
import threading as _threading

import numpy as np
import torch

from . import _lib
from . import batch as _batch
from ._lib import CLASS_64, CLASS_192, CLASS_VAR
from ._runtime import MSG_TOO_LONG, bad_base_message

MIN_64_NT, MAX_64_NT = 0, 32
MIN_192_NT, MAX_192_NT = 33, 96
MIN_VAR_NT, MAX_VAR_NT = 97, 1024
MAX_REPR_LEN = 75
_M64 = (1 << 64) - 1


def get_domain_64():
    return MIN_64_NT, MAX_64_NT


def get_domain_192():
    return MIN_192_NT, MAX_192_NT


def get_domain_var():
    return MIN_VAR_NT, MAX_VAR_NT


def _nblocks(length):
    return (length + 31) // 32


class _One:
    """Operand buffers of the single-object calls (one set per process; the lock keeps two Python threads from
    interleaving on them -- ctypes releases the GIL during a call)."""

    def __init__(self):
        self.ctx = _batch.context()
        self.lib = _lib.lib()
        self.wa = (_C.c_uint64 * 32)()
        self.wb = (_C.c_uint64 * 32)()
        self.klass = _C.c_int32()
        self.bad = _C.c_int32()
        self.text = _C.create_string_buffer(1024)
        self.lock = _threading.Lock()


_ONE = None


def _one():
    global _ONE
    if _ONE is None:
        _ONE = _One()
    return _ONE


class _ShortSeqBase:
    """Shared protocol.  `_packed` is a tuple of Python ints (uint64 blocks), `_length` an int, `_hash` the
    precomputed `hash()` (the reference also hashes in O(1): it returns the first block)."""

    __slots__ = ("_packed", "_length", "_hash")
    _klass = None

    def __init__(self, *a, **k):
        raise TypeError(f"{type(self).__name__} objects are created with shortseq_b200.pack()")

    def __hash__(self):
        # prehash = first block, seen as Py_hash_t (reference short_seq_64.pyx:35-36; -1 -> -2 by CPython)
        return self._hash

    def __len__(self):
        return self._length

    def __eq__(self, other):
        if type(other) is type(self):
            nb = _nblocks(self._length)
            return self._length == other._length and self._packed[:nb] == other._packed[:nb]
        if isinstance(other, (str, bytes)):
            return self._length == len(other) and str(self) == other   # bytes never compare equal
        return False

    def __ne__(self, other):
        return not self.__eq__(other)

    def __getitem__(self, item):
        if isinstance(item, slice):
            start, stop, step = item.indices(self._length)
            if step != 1:
                raise TypeError("Slice step not supported")
            slice_len = max(0, stop - start)
            if slice_len == 0:
                return empty
            if slice_len == 1:
                return _subscript(self._packed, start)
            return _slice(self._packed, start, slice_len)
        if isinstance(item, int):
            index = item
            if index < 0:
                index += self._length
            if index < 0 or index >= self._length:
                raise IndexError("Sequence index out of range")
            return _subscript(self._packed, index)
        raise TypeError(f"Invalid index type: {type(item)}")

    def __xor__(self, other):
        if type(other) is not type(self):
            raise TypeError(f"Argument 'other' has incorrect type (expected {type(self).__name__}, "
                            f"got {type(other).__name__})")
        if self._length != other._length:
            raise Exception(f"Hamming distance requires sequences of equal length "
                            f"({self._length} != {other._length})")
        if self._length == 0:
            return 0
        o = _one()
        with o.lock:
            for i, w in enumerate(self._packed):
                o.wa[i] = w
            for i, w in enumerate(other._packed):
                o.wb[i] = w
            _lib.check(o.lib.ssq_hamming_one(o.ctx.handle, o.wa, o.wb, self._length, _C.byref(o.klass)))
            return o.klass.value

    def __str__(self):
        if self._length == 0:
            return ""
        o = _one()
        with o.lock:
            for i, w in enumerate(self._packed):
                o.wa[i] = w
            _lib.check(o.lib.ssq_decode_one(o.ctx.handle, o.wa, self._length, o.text))
            return o.text.raw[:self._length].decode("ascii")

    def __repr__(self):
        return f"<{type(self).__name__} ({self._length} nt): {self}>"

    def __reduce__(self):
        return (_box, (self._klass, self._packed, self._length))


class ShortSeq64(_ShortSeqBase):
    """0..32 nt in one 64-bit block (reference short_seq_64.pxd:11-14)."""
    __slots__ = ()
    _klass = CLASS_64


class ShortSeq192(_ShortSeqBase):
    """33..96 nt in three 64-bit blocks, unused blocks zero (reference short_seq_192.pxd:11-14)."""
    __slots__ = ()
    _klass = CLASS_192


class ShortSeqVar(_ShortSeqBase):
    """97..1024 nt in ceil(len/32) blocks (reference short_seq_var.pxd:14-17)."""
    __slots__ = ()
    _klass = CLASS_VAR

    def __sizeof__(self):
        # reference short_seq_var.pyx:83-84: sizeof(ShortSeqVar) + 8 per block
        return 32 + 8 * _nblocks(self._length)

    def __repr__(self):
        # reference short_seq_var.pyx:86-89: at most MAX_REPR_LEN characters
        return f"<ShortSeqVar ({self._length} nt): {str(self)[:MAX_REPR_LEN]} ... >"


_TYPES = {CLASS_64: ShortSeq64, CLASS_192: ShortSeq192, CLASS_VAR: ShortSeqVar}


def _hash_of_block(h):
    if h >= 1 << 63:
        h -= 1 << 64
    return -2 if h == -1 else h


def _box(klass, packed, length):
    """Build a ShortSeq object from its class, blocks and length (no validation)."""
    cls = _TYPES[klass]
    obj = object.__new__(cls)
    obj._packed = tuple(int(x) & _M64 for x in packed)
    obj._length = int(length)
    obj._hash = _hash_of_block(obj._packed[0])
    return obj


class _NoGC:
    """Bulk creation of millions of small objects: the cyclic collector would rescan the growing heap over and over
    (3x the time of the loop itself); none of these objects can be part of a cycle."""

    def __enter__(self):
        import gc
        self._was = gc.isenabled()
        gc.disable()

    def __exit__(self, *exc):
        if self._was:
            import gc
            gc.enable()
        return False


def _box_many(klass, words, lens):
    """Box a whole array of fixed-class keys: words uint64 ndarray [n] (ShortSeq64) or [n, 3] (ShortSeq192), lens
    ndarray [n] -> list of ShortSeq objects.  The per-object work is three slot stores; blocks, lengths and hashes are
    converted to Python ints in bulk."""
    cls = _TYPES[klass]
    w = np.ascontiguousarray(words).view(np.uint64)
    first = w if w.ndim == 1 else w[:, 0]
    hashes = first.view(np.int64).copy()
    hashes[hashes == -1] = -2
    with _NoGC():
        packed = [(x,) for x in w.tolist()] if w.ndim == 1 else [tuple(r) for r in w.tolist()]
    new = object.__new__
    out = []
    append = out.append
    with _NoGC():
        for p, l, h in zip(packed, np.asarray(lens).tolist(), hashes.tolist()):
            o = new(cls)
            o._packed = p
            o._length = l
            o._hash = h
            append(o)
    return out


empty = _box(CLASS_64, (0,), 0)   # module singleton, like the reference's (short_seq.pyx:7)


def _as_arrays(objs):
    """ShortSeq objects (any mix of classes) -> [(klass, positions, ShortSeqArray)] on the current GPU, one per class."""
    ctx = _batch.context()
    groups = {}
    for i, o in enumerate(objs):
        groups.setdefault(o._klass, []).append(i)
    out = []
    for k, idx in groups.items():
        sel = [objs[i] for i in idx]
        lens = np.fromiter((o._length for o in sel), dtype=np.int64, count=len(sel))
        if k == CLASS_VAR:
            nb = (lens + 31) // 32
            wo = np.zeros(len(sel) + 1, dtype=np.int64)
            np.cumsum(nb, out=wo[1:])
            flat = np.fromiter((w for o in sel for w in o._packed), dtype=np.uint64, count=int(wo[-1]))
            arr = _batch.ShortSeqArray(ctx, k, torch.from_numpy(flat.view(np.int64)).to(ctx.device),
                                       torch.from_numpy(lens.astype(np.int16)).to(ctx.device), torch.from_numpy(wo).to(ctx.device))
        else:
            W = 1 if k == CLASS_64 else 3
            flat = np.fromiter((w for o in sel for w in o._packed), dtype=np.uint64, count=W * len(sel)).view(np.int64)
            words = torch.from_numpy(flat if W == 1 else flat.reshape(-1, 3)).to(ctx.device)
            arr = _batch.ShortSeqArray(ctx, k, words, torch.from_numpy(lens.astype(np.uint8)).to(ctx.device))
        out.append((k, idx, arr))
    return out


def decode_many(seqs):
    """str() of many ShortSeq objects with one decode kernel per class instead of one launch per object
    (reference __str__: short_seq_64.pyx:114-121 etc.) -> list of str in the order given."""
    seqs = list(seqs)
    out = [""] * len(seqs)
    for _, idx, arr in _as_arrays([s for s in seqs]):
        for i, text in zip(idx, arr.decode_to_list()):
            out[i] = text
    return out


def hamming_many(a, b):
    """a[i] ^ b[i] for two equal-length lists of ShortSeq objects, one kernel per class -> list of int.
    Same rules as the reference's __xor__: same class and equal lengths pairwise."""
    a, b = list(a), list(b)
    if len(a) != len(b):
        raise ValueError("hamming_many needs two lists of the same length")
    out = [0] * len(a)
    for x, y in zip(a, b):
        if type(x) is not type(y):
            raise TypeError(f"Argument 'other' has incorrect type (expected {type(x).__name__}, got {type(y).__name__})")
    ga, gb = _as_arrays(a), _as_arrays(b)
    for (_, idx, arr_a), (_, _, arr_b) in zip(ga, gb):
        d = _batch.hamming_batch(arr_a, arr_b).cpu().numpy()
        for i, v in zip(idx, d.tolist()):
            out[i] = int(v)
    return out


def _as_array(obj):
    """A one-element ShortSeqArray on the current GPU holding this object."""
    ctx = _batch.context()
    k = obj._klass
    words = np.array(obj._packed, dtype=np.uint64).view(np.int64)
    if k == CLASS_64:
        w = torch.from_numpy(words.reshape(1)).to(ctx.device)
        return _batch.ShortSeqArray(ctx, k, w, torch.tensor([obj._length], dtype=torch.uint8, device=ctx.device))
    if k == CLASS_192:
        w = torch.from_numpy(words.reshape(1, 3)).to(ctx.device)
        return _batch.ShortSeqArray(ctx, k, w, torch.tensor([obj._length], dtype=torch.uint8, device=ctx.device))
    w = torch.from_numpy(words).to(ctx.device)
    wo = torch.tensor([0, len(obj._packed)], dtype=torch.int64, device=ctx.device)
    return _batch.ShortSeqArray(ctx, k, w, torch.tensor([obj._length], dtype=torch.int16, device=ctx.device), wo)


# === constructors (reference short_seq.pyx:13-74) ===============================================

def _new(data: bytes):
    length = len(data)
    if length == 0:
        return empty
    if length > MAX_VAR_NT:
        raise Exception(MSG_TOO_LONG)
    o = _one()
    with o.lock:
        rc = o.lib.ssq_pack_one(o.ctx.handle, data, length, o.wa, _C.byref(o.klass), _C.byref(o.bad))
        if rc == _lib.ERR_BAD_BASE:
            raise Exception(bad_base_message(data))
        _lib.check(rc)
        k = o.klass.value
        return _box(k, o.wa[:1 if k == CLASS_64 else (3 if k == CLASS_192 else _nblocks(length))], length)


def pack(seq, /):
    """Pack a str / bytes into the ShortSeq class its length selects (reference short_seq.pyx:13-28)."""
    if isinstance(seq, str):
        return empty if not seq else _new(seq.encode("latin-1", errors="replace"))
    if isinstance(seq, bytes):
        return empty if not seq else _new(seq)
    if type(seq) in (ShortSeq64, ShortSeq192, ShortSeqVar):
        return seq
    raise TypeError(f'Cannot pack objects of type "{type(seq)}"')


def from_str(seq_str):
    if not isinstance(seq_str, str):
        raise TypeError(f"Argument 'seq_str' has incorrect type (expected str, got {type(seq_str).__name__})")
    return empty if not seq_str else _new(seq_str.encode("latin-1", errors="replace"))


def from_bytes(seq_bytes):
    if not isinstance(seq_bytes, bytes):
        raise TypeError(f"Argument 'seq_bytes' has incorrect type (expected bytes, got {type(seq_bytes).__name__})")
    return empty if not seq_bytes else _new(seq_bytes)


# === subscript / slice (reference short_seq.pyx:78-238) ==========================================

def _subscript(packed, index):
    """One base as a ShortSeq64 (reference short_seq.pyx:78-91)."""
    return _box(CLASS_64, ((packed[index // 32] >> (2 * (index % 32))) & 3,), 1)


def _slice(packed, start, slice_len):
    """Sub-sequence [start, start+slice_len); the class of the result follows its length
    (reference short_seq.pyx:94-116); blocks are funnel-shifted and the tail trimmed
    (short_seq.pyx:202-238) so whole-block compares / popcounts stay valid."""
    big = 0
    for i, w in enumerate(packed):
        big |= w << (64 * i)
    big = (big >> (2 * start)) & ((1 << (2 * slice_len)) - 1)
    if slice_len <= MAX_64_NT:
        klass, nb = CLASS_64, 1
    elif slice_len <= MAX_192_NT:
        klass, nb = CLASS_192, 3
    else:
        klass, nb = CLASS_VAR, _nblocks(slice_len)
    return _box(klass, tuple((big >> (64 * i)) & _M64 for i in range(nb)), slice_len)
