"""Batch entry points: pack / decode / Hamming over whole arrays of reads on the GPU.

These are the drop-ins for the reference's per-object paths (`sq.pack`, `str()`,
`a ^ b`): one call handles a whole batch held as a contiguous ASCII buffer plus
offsets, and returns an array-backed result (`ShortSeqArray`) whose items box
into ShortSeq64 / ShortSeq192 / ShortSeqVar objects on demand.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import CLASS_64, CLASS_192, CLASS_VAR
from ._runtime import (MSG_TOO_LONG, ShortSeqClassError, bad_base_message, context, gather_reads, ptr, to_device,
                       words_to_numpy)

MIN_NT = {CLASS_64: 0, CLASS_192: 33, CLASS_VAR: 97}
MAX_NT = {CLASS_64: 32, CLASS_192: 96, CLASS_VAR: 1024}
WORDS = {CLASS_64: 1, CLASS_192: 3}


def class_of_length(n):
    """Container class by length (reference short_seq.pyx:54-74)."""
    if n <= 32:
        return CLASS_64
    if n <= 96:
        return CLASS_192
    if n <= 1024:
        return CLASS_VAR
    raise Exception(MSG_TOO_LONG)


class ReadBatch:
    """Reads as one ASCII buffer + offsets, resident on a GPU (host copy kept when there is one)."""

    def __init__(self, ctx, ascii_t, offsets_t, host_ascii=None, host_offsets=None):
        self.ctx, self.ascii, self.offsets = ctx, ascii_t, offsets_t
        self.host_ascii, self.host_offsets = host_ascii, host_offsets
        self.n = int(offsets_t.numel()) - 1

    @classmethod
    def make(cls, source, offsets=None, device=None):
        if isinstance(source, ReadBatch):
            return source
        ctx = context(device if not isinstance(source, torch.Tensor) else source.device)
        if offsets is None:
            if isinstance(source, (list, tuple)):
                source = [s.encode("latin-1") if isinstance(s, str) else s for s in source]
                h_ascii, h_off = gather_reads(list(source))
            else:
                raise TypeError("pass a list of bytes, or an ASCII buffer together with offsets")
            return cls(ctx, to_device(ctx, h_ascii, torch.uint8), to_device(ctx, h_off, torch.int64), h_ascii, h_off)
        if isinstance(source, torch.Tensor):
            a = source.to(torch.uint8).contiguous()
            o = offsets if isinstance(offsets, torch.Tensor) else torch.as_tensor(np.asarray(offsets, dtype=np.int64))
            return cls(ctx, a, o.to(ctx.device, torch.int64).contiguous())
        if isinstance(source, (bytes, bytearray, memoryview)):
            source = np.frombuffer(source, dtype=np.uint8)
        h_ascii = np.ascontiguousarray(source, dtype=np.uint8)
        h_off = np.ascontiguousarray(offsets, dtype=np.int64)
        return cls(ctx, to_device(ctx, h_ascii, torch.uint8), to_device(ctx, h_off, torch.int64), h_ascii, h_off)

    def read_bytes(self, i):
        """Bytes of read i (small device->host copy when no host copy is held)."""
        if self.host_ascii is not None:
            return bytes(self.host_ascii[self.host_offsets[i]: self.host_offsets[i + 1]])
        o = self.offsets[i: i + 2].cpu()
        return bytes(self.ascii[int(o[0]): int(o[1])].cpu().numpy())

    def lengths_host(self):
        if self.host_offsets is not None:
            return np.diff(self.host_offsets)
        return np.diff(self.offsets.cpu().numpy())


def raise_for_report(rep, batch=None, index_map=None):
    """Turn a data error reported by the device into the exception the reference raises."""
    if rep.code == _lib.OK:
        return
    i = int(rep.first_bad_read)
    if rep.code == _lib.ERR_BAD_BASE:
        msg = bad_base_message(batch.read_bytes(i)) if batch is not None else "Unsupported base character"
        raise Exception(msg)
    if rep.code == _lib.ERR_TOO_LONG:
        raise Exception(MSG_TOO_LONG)
    where = int(index_map[i]) if index_map is not None else i
    if rep.code == _lib.ERR_CLASS:
        raise ShortSeqClassError(f"read {where}: length outside the container class of this batch call")
    if rep.code == _lib.ERR_LEN_MISMATCH:
        raise Exception(f"Hamming distance requires sequences of equal length (pair {where})")
    if rep.code == _lib.ERR_TABLE_FULL:
        raise _lib.LibraryError("counter table overflow")
    if rep.code == _lib.ERR_EXCHANGE:
        raise _lib.LibraryError("multi-GPU merge: another rank's share never arrived")
    raise _lib.LibraryError(f"device reported status {rep.code}")


class ShortSeqArray:
    """A packed batch of one container class, resident on the GPU.

    words: int64 tensor holding the uint64 bit patterns -- [n] (ShortSeq64), [n, 3]
    (ShortSeq192) or flat CSR (ShortSeqVar, with word_off[n+1]); lens: uint8 ([n]) or
    int16 for ShortSeqVar.
    """

    def __init__(self, ctx, klass, words, lens, word_off=None):
        self.ctx, self.klass, self.words, self.lens, self.word_off = ctx, klass, words, lens, word_off
        self._host = None

    def __len__(self):
        return int(self.lens.numel())

    # -- host views -----------------------------------------------------------------------
    def to_host(self):
        """(words uint64, lens, word_off or None) as numpy arrays."""
        if self._host is None:
            w = words_to_numpy(self.words)
            l = self.lens.cpu().numpy()
            wo = self.word_off.cpu().numpy() if self.word_off is not None else None
            self._host = (w, l, wo)
        return self._host

    def prehash(self):
        """Raw u64 prehash per read = first packed word (reference __hash__, short_seq_64.pyx:35-36)."""
        w, _, wo = self.to_host()
        if self.klass == CLASS_64:
            return w.copy()
        if self.klass == CLASS_192:
            return w[:, 0].copy()
        return w[wo[:-1]].copy()

    def __getitem__(self, i):
        """Box read i into the matching ShortSeq object."""
        from .short_seq import _box
        w, l, wo = self.to_host()
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError("ShortSeqArray index out of range")
        length = int(l[i])
        if self.klass == CLASS_64:
            blocks = (int(w[i]),)
        elif self.klass == CLASS_192:
            blocks = tuple(int(x) for x in w[i])
        else:
            blocks = tuple(int(x) for x in w[wo[i]: wo[i + 1]])
        # the object class follows the element's own length (an element of a slice_batch result may be shorter than
        # its array's class, reference short_seq.pyx:94-116); bits >= 2 len are zero, so dropping / padding blocks is exact
        k = class_of_length(length)
        nb = 1 if k == CLASS_64 else (3 if k == CLASS_192 else max(1, (length + 31) // 32))
        blocks = (blocks + (0,) * nb)[:nb]
        return _box(k, blocks, length)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    # -- device ops ---------------------------------------------------------------------------
    def decode(self):
        """-> (ascii uint8 tensor, offsets int64 tensor [n+1]) on the GPU."""
        return decode_batch(self)

    def decode_to_list(self):
        a, o = self.decode()
        a, o = a.cpu().numpy().tobytes(), o.cpu().numpy()
        return [a[o[i]: o[i + 1]].decode("ascii") for i in range(len(self))]

    def hamming(self, other):
        return hamming_batch(self, other)

    def slice(self, start=0, stop=None):
        return slice_batch(self, start, stop)

    def kmers(self, k, stride=1):
        return kmers_batch(self, k, stride)


def _alloc_out(ctx, klass, n):
    if klass == CLASS_64:
        return ctx.empty((n,), torch.int64), ctx.empty((n,), torch.uint8)
    return ctx.empty((n, 3), torch.int64), ctx.empty((n,), torch.uint8)


def _pack_raw(b, klass):
    """Enqueue the pack kernel of one class and synchronise -> (ShortSeqArray, Report)."""
    ctx = b.ctx
    L = _lib.lib()
    h = ctx.bind()
    n, nbytes = b.n, int(b.ascii.numel())
    if klass in (CLASS_64, CLASS_192):
        words, lens = _alloc_out(ctx, klass, n)
        fn = L.ssq_pack64 if klass == CLASS_64 else L.ssq_pack192
        _lib.check(fn(h, ptr(b.ascii), nbytes, ptr(b.offsets), n, ptr(words), ptr(lens)))
        word_off = None
    elif klass == CLASS_VAR:
        bound = L.ssq_packvar_words_bound(nbytes, n)
        words = ctx.empty((bound,), torch.int64)
        lens = ctx.empty((n,), torch.int16)
        word_off = ctx.empty((n + 1,), torch.int64)
        _lib.check(L.ssq_packvar(h, ptr(b.ascii), nbytes, ptr(b.offsets), n, ptr(word_off), ptr(words), ptr(lens)))
    else:
        raise ValueError("klass must be CLASS_64, CLASS_192 or CLASS_VAR")
    rep = ctx.sync()
    if klass == CLASS_VAR and rep.code == _lib.OK:
        words = words[: int(word_off[-1])]
    return ShortSeqArray(ctx, klass, words, lens, word_off), rep


def pack_batch(source, offsets=None, klass=None, device=None):
    """Pack a class-homogeneous batch of reads on the GPU -> ShortSeqArray.

    source: list of bytes, or an ASCII buffer (bytes / numpy uint8 / CUDA uint8 tensor) with
    `offsets` (n+1 int64).  klass: CLASS_64 / CLASS_192 / CLASS_VAR; inferred from the first
    read when omitted.  Raises what the reference's sq.pack raises for the first offending read
    (in list order); a read of another length class raises ShortSeqClassError -- use
    pack_mixed() for mixed batches.
    """
    b = ReadBatch.make(source, offsets, device)
    if klass is None:
        if b.n == 0:
            klass = CLASS_64
        else:
            first = b.offsets[:2].cpu() if b.host_offsets is None else b.host_offsets[:2]
            klass = class_of_length(int(first[1]) - int(first[0]))
    arr, rep = _pack_raw(b, klass)
    raise_for_report(rep, b)
    return arr


def normalize_batch(source, offsets=None, device=None):
    """Opt-in tolerant alphabet: a ReadBatch in which a c g t are upper-cased and u / U read as T.
    The reference rejects all of these (its table_91 maps U to T's code but the validator refuses it, util.pyx:44-50);
    pack_batch(normalize_batch(...)) accepts them."""
    b = ReadBatch.make(source, offsets, device)
    out = torch.empty_like(b.ascii)
    _lib.check(_lib.lib().ssq_normalize(b.ctx.bind(), ptr(b.ascii), int(b.ascii.numel()), ptr(out)))
    return ReadBatch(b.ctx, out, b.offsets)


def slice_batch(arr, start=0, stop=None):
    """arr[i][start:stop] for every read of a ShortSeqArray in one launch (reference short_seq.pyx:94-238, per object).

    start / stop: Python ints with slice semantics (None, negative values), or int64 tensors / arrays with one value per
    read (non-negative).  The class of the result follows the largest possible slice length, as the reference's result
    class follows the slice length: <= 32 -> CLASS_64, <= 96 -> CLASS_192, else CLASS_VAR.  Boxing an element of the
    result (`out[i]`) picks the object class from that element's own length."""
    ctx, L, n = arr.ctx, _lib.lib(), len(arr)
    max_len = MAX_NT[arr.klass]

    def per_read(v):
        return v is not None and not isinstance(v, int)

    starts = stops = None
    start0, stop0 = 0, max_len
    if per_read(start) or per_read(stop) or (isinstance(start, int) and start < 0) or (isinstance(stop, int) and stop < 0):
        lens = arr.lens.to(torch.int64)
        st = torch.zeros_like(lens) if start is None else (to_device(ctx, start, torch.int64) if per_read(start) else torch.full_like(lens, start))
        sp = lens.clone() if stop is None else (to_device(ctx, stop, torch.int64) if per_read(stop) else torch.full_like(lens, stop))
        st = torch.where(st < 0, torch.clamp(st + lens, min=0), st)
        sp = torch.where(sp < 0, torch.clamp(sp + lens, min=0), sp)
        starts, stops = st.contiguous(), sp.contiguous()
        width = int(torch.clamp(torch.minimum(sp, lens) - torch.minimum(st, lens), min=0).max()) if n else 0
    else:
        start0 = 0 if start is None else start
        stop0 = max_len if stop is None else stop
        width = max(0, min(stop0, max_len) - start0)
    out_klass = CLASS_64 if width <= 32 else (CLASS_192 if width <= 96 else CLASS_VAR)
    h = ctx.bind()
    word_off = None
    if out_klass == CLASS_VAR:
        word_off = ctx.empty((n + 1,), torch.int64)
        _lib.check(L.ssq_slice_words(h, arr.klass, ptr(arr.lens), n, ptr(starts), ptr(stops), start0, stop0, width, ptr(word_off)))
        words = ctx.empty((int(word_off[-1]) if n else 0,), torch.int64)
        lens_out = ctx.empty((n,), torch.int16)
    else:
        words, lens_out = _alloc_out(ctx, out_klass, n)
    _lib.check(L.ssq_slice(h, arr.klass, ptr(arr.words), ptr(arr.word_off), ptr(arr.lens), n, ptr(starts), ptr(stops), start0, stop0,
                           width, out_klass, ptr(words), ptr(word_off), ptr(lens_out)))
    raise_for_report(ctx.sync())
    return ShortSeqArray(ctx, out_klass, words, lens_out, word_off)


def kmers_batch(arr, k, stride=1):
    """Every k-mer (k <= 32) of every read -> (ShortSeqArray of CLASS_64 k-mers grouped by read, kmer_off int64 [n+1]).
    `DeviceCounter(CLASS_64).insert(kmers)` then counts them."""
    if not 1 <= k <= 32 or stride < 1:
        raise ValueError("k must be 1..32 and stride >= 1")
    ctx, L, n = arr.ctx, _lib.lib(), len(arr)
    h = ctx.bind()
    kmer_off = ctx.empty((n + 1,), torch.int64)
    _lib.check(L.ssq_kmers_count(h, arr.klass, ptr(arr.lens), n, k, stride, ptr(kmer_off)))
    total = int(kmer_off[-1]) if n else 0
    words, lens = ctx.empty((total,), torch.int64), ctx.empty((total,), torch.uint8)
    _lib.check(L.ssq_kmers64(h, arr.klass, ptr(arr.words), ptr(arr.word_off), ptr(arr.lens), n, k, stride, ptr(kmer_off), ptr(words), ptr(lens)))
    raise_for_report(ctx.sync())
    return ShortSeqArray(ctx, CLASS_64, words, lens), kmer_off


def umi_collapse(umis, counts, group_off=None, threshold=1, method="directional"):
    """Cluster distinct UMIs by Hamming distance with UMI-tools' rules (README.md:82-88 of the reference).

    umis: ShortSeqArray of CLASS_64 (distinct UMIs, e.g. the keys of DeviceCounter.export); counts: their multiplicities
    (int64 tensor / array); group_off: int64 [g + 1] boundaries of independent groups (None: one group).
    method "directional" (count[a] >= 2 count[b] - 1 along every edge) or "cluster" (connected components).
    -> (rep int64 tensor [n]: index of every UMI's representative, cluster_counts int64 tensor [n]: summed counts at the
    representatives and 0 elsewhere, n_clusters int64 tensor [g])."""
    if umis.klass != CLASS_64:
        raise TypeError("umi_collapse needs a ShortSeq64 array (UMIs are at most 32 nt)")
    ctx, n = umis.ctx, len(umis)
    cnt = to_device(ctx, counts, torch.int64)
    if cnt.numel() != n:
        raise ValueError("counts and umis differ in size")
    goff = torch.tensor([0, n], dtype=torch.int64, device=ctx.device) if group_off is None else to_device(ctx, group_off, torch.int64)
    g = int(goff.numel()) - 1
    rep = ctx.empty((n,), torch.int64)
    ncl = ctx.empty((g,), torch.int64)
    m = {"directional": 0, "cluster": 1}[method]
    _lib.check(_lib.lib().ssq_umi_cluster(ctx.bind(), ptr(umis.words), ptr(umis.lens), ptr(cnt), n, ptr(goff), g, int(threshold), m,
                                          ptr(rep), ptr(ncl)))
    raise_for_report(ctx.sync())
    cluster_counts = torch.zeros_like(cnt).index_add_(0, rep, cnt)
    return rep, cluster_counts, ncl


def split_by_class(h_ascii, h_off):
    """Host-side split of a mixed batch into per-class sub-batches.

    -> list of (klass, original_positions, sub_ascii, sub_offsets); reads longer than 1024 are
    left out (the caller reports them).
    """
    lens = np.diff(h_off)
    klass_of = np.where(lens <= 32, CLASS_64, np.where(lens <= 96, CLASS_192, np.where(lens <= 1024, CLASS_VAR, -1)))
    out = []
    for k in (CLASS_64, CLASS_192, CLASS_VAR):
        idx = np.nonzero(klass_of == k)[0]
        if idx.size == 0:
            continue
        sub_off = np.zeros(idx.size + 1, np.int64)
        np.cumsum(lens[idx], out=sub_off[1:])
        total = int(sub_off[-1])
        if idx.size == lens.size:
            sub_ascii = h_ascii
        else:
            take = np.repeat(h_off[idx] - sub_off[:-1], lens[idx]) + np.arange(total, dtype=np.int64)
            sub_ascii = h_ascii[take]
        out.append((k, idx, sub_ascii, sub_off))
    return out


def pack_mixed(reads, device=None):
    """Pack a list of bytes of any lengths: reads are split by container class on the host and
    each class is packed by its kernel.  -> ({klass: ShortSeqArray}, {klass: original positions}).
    Like a loop over the reference's sq.pack, the first offending read in list order decides the
    exception."""
    h_ascii, h_off = gather_reads([r.encode("latin-1") if isinstance(r, str) else r for r in reads])
    lens = np.diff(h_off)
    errors = []
    too_long = np.nonzero(lens > 1024)[0]
    if too_long.size:
        errors.append((int(too_long[0]), Exception(MSG_TOO_LONG)))
    arrays, index = {}, {}
    for k, idx, sub_ascii, sub_off in split_by_class(h_ascii, h_off):
        b = ReadBatch.make(sub_ascii, sub_off, device)
        arr, rep = _pack_raw(b, k)
        if rep.code != _lib.OK:
            try:
                raise_for_report(rep, b, idx)
            except Exception as e:  # noqa: BLE001 -- re-raised below in list order
                errors.append((int(idx[int(rep.first_bad_read)]), e))
        arrays[k], index[k] = arr, idx
    if errors:
        raise min(errors, key=lambda t: t[0])[1]
    return arrays, index


def decode_batch(arr):
    """ShortSeqArray -> (ascii uint8 tensor, offsets int64 tensor[n+1]), both on the GPU."""
    ctx = arr.ctx
    L = _lib.lib()
    h = ctx.bind()
    n = len(arr)
    out_off = ctx.empty((n + 1,), torch.int64)
    if arr.klass in (CLASS_64, CLASS_192) and n > 0:
        # fused offsets: per-tile totals + a scan over n/512 entries; the decode kernel writes the offsets itself
        ntiles = (n + 511) // 512
        tile_base = ctx.empty((2 * ntiles + 2,), torch.int64)
        _lib.check(L.ssq_decode_tiles(h, ptr(arr.lens), n, 32 if arr.klass == CLASS_64 else 96, ptr(tile_base)))
        total = int(tile_base[ntiles])
        out = ctx.empty((max(total, 1),), torch.uint8)
        fn = L.ssq_decode64_fused if arr.klass == CLASS_64 else L.ssq_decode192_fused
        _lib.check(fn(h, ptr(arr.words), ptr(arr.lens), n, ptr(tile_base), ptr(out_off), ptr(out)))
        raise_for_report(ctx.sync())
        return out[:total], out_off
    _lib.check(L.ssq_lens_to_offsets(h, ptr(arr.lens), 2 if arr.klass == CLASS_VAR else 1, n, ptr(out_off)))
    total = int(out_off[-1]) if n else 0
    out = ctx.empty((max(total, 1),), torch.uint8)
    if arr.klass == CLASS_64:
        _lib.check(L.ssq_decode64(h, ptr(arr.words), ptr(arr.lens), n, ptr(out_off), ptr(out)))
    elif arr.klass == CLASS_192:
        _lib.check(L.ssq_decode192(h, ptr(arr.words), ptr(arr.lens), n, ptr(out_off), ptr(out)))
    else:
        _lib.check(L.ssq_decodevar(h, ptr(arr.words), ptr(arr.word_off), ptr(arr.lens), n, ptr(out_off), ptr(out)))
    raise_for_report(ctx.sync())
    return out[:total], out_off


def hamming_batch(a, b):
    """Element-wise Hamming distance of two ShortSeqArrays of the same class (`a[i] ^ b[i]`).

    -> uint8 tensor (int16 for ShortSeqVar).  Like the reference, pairs of different length
    raise ("Hamming distance requires sequences of equal length").
    """
    if a.klass != b.klass:
        raise TypeError("Hamming distance requires arrays of the same ShortSeq class")
    if len(a) != len(b):
        raise ValueError("arrays differ in size")
    ctx = a.ctx
    L = _lib.lib()
    h = ctx.bind()
    n = len(a)
    if a.klass == CLASS_VAR:
        dist = ctx.empty((n,), torch.int16)
        _lib.check(L.ssq_hamming_pairsvar(h, ptr(a.words), ptr(a.word_off), ptr(a.lens), ptr(b.words), ptr(b.word_off),
                                          ptr(b.lens), n, ptr(dist)))
    else:
        dist = ctx.empty((n,), torch.uint8)
        fn = L.ssq_hamming_pairs64 if a.klass == CLASS_64 else L.ssq_hamming_pairs192
        _lib.check(fn(h, ptr(a.words), ptr(a.lens), ptr(b.words), ptr(b.lens), n, ptr(dist)))
    raise_for_report(ctx.sync())
    return dist


def hamming_refset(queries, refs, thresh=1):
    """Each query against a reference set (UMI-collapse style).

    -> (min_dist uint8, argmin int64 (-1 when no ref has the query's length), n_within int32):
    nearest reference per query and the number of references within `thresh`.
    """
    if queries.klass != refs.klass or queries.klass == CLASS_VAR:
        raise TypeError("hamming_refset needs two ShortSeq64 or two ShortSeq192 arrays")
    ctx = queries.ctx
    L = _lib.lib()
    h = ctx.bind()
    nq, nr = len(queries), len(refs)
    min_dist = ctx.empty((nq,), torch.uint8)
    argmin = ctx.empty((nq,), torch.int32)
    within = ctx.empty((nq,), torch.int32)
    _lib.check(L.ssq_hamming_refset(h, WORDS[queries.klass], ptr(queries.words), ptr(queries.lens), nq, ptr(refs.words),
                                    ptr(refs.lens), nr, int(thresh), ptr(min_dist), ptr(argmin), ptr(within)))
    raise_for_report(ctx.sync())
    return min_dist, argmin.to(torch.int64), within


def synth_reads(n, n_keys, len_lo, len_hi, seed=0x5EED0001, first_read=0, device=None):
    """Deterministic synthetic batch generated on the GPU (same bytes as oracle.synth_reads)."""
    ctx = context(device)
    L = _lib.lib()
    h = ctx.bind()
    offsets = ctx.empty((n + 1,), torch.int64)
    ascii_t = ctx.empty((max(1, n * len_hi),), torch.uint8)
    _lib.check(L.ssq_synth_reads(h, C.c_uint64(seed), first_read, n, n_keys, len_lo, len_hi, ptr(offsets), ptr(ascii_t)))
    raise_for_report(ctx.sync())
    total = int(offsets[-1]) if n else 0
    return ReadBatch(ctx, ascii_t[:total] if len_lo != len_hi else ascii_t[: n * len_hi], offsets)
