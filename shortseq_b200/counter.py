"""Dedup counting: the drop-in ShortSeqCounter and the device-resident DeviceCounter behind it.

Reference: shortseq/counter.pyx:10-71.  The reference fills a Python dict serially
(pack, dict probe, PyLong increment per read); here the whole list goes through the fused
pack+count kernel and only the uniques are boxed into ShortSeq objects.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import batch as _batch
from ._lib import CLASS_64, CLASS_192, CLASS_VAR
from ._runtime import MSG_TOO_LONG, bad_base_message, context, gather_reads, ptr, words_to_numpy
from .short_seq import ShortSeq64, ShortSeq192, ShortSeqVar, _NoGC, _box, _box_many


class DeviceCounter:
    """A GPU hash table of (length, packed words) -> count for one container class.

    klass: CLASS_64 or CLASS_192 (the reference never deduplicates ShortSeqVar keys,
    SURVEY trap T3).  expected_unique (a bound on the distinct keys; 0 = unknown) sizes the table to at least twice
    the bound and lets a batch be counted in one pass.  The table grows before a pass that could load it beyond 75 %
    and after a pass that left it more than 60 % full; once it holds more keys than the bound it continues in the
    conservative mode of expected_unique = 0 (gated sub-batches, grows as needed).  Only one single pass that brings
    far more distinct keys than the bound can still overflow (LibraryError "counter table overflow").
    """

    def __init__(self, klass, expected_unique=0, hash_rot=0, device=None):
        if klass not in (CLASS_64, CLASS_192):
            raise ValueError("DeviceCounter supports CLASS_64 and CLASS_192")
        self.ctx = context(device)
        self.klass = klass
        self.W = 1 if klass == CLASS_64 else 3
        h = C.c_void_p()
        _lib.check(_lib.lib().ssq_counter_create(self.ctx.bind(), klass, int(expected_unique), int(hash_rot), C.byref(h)))
        self.handle = h

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            try:
                _lib.lib().ssq_counter_destroy(h)
            except Exception:  # noqa: BLE001 -- interpreter shutdown
                pass
            self.handle = None

    # -- updates ------------------------------------------------------------------------------
    def pack_count(self, source, offsets=None, check=True):
        """Fused pack + count of a batch of reads of this counter's class -> ShortSeqArray."""
        b = _batch.ReadBatch.make(source, offsets, self.ctx.device)
        words, lens = _batch._alloc_out(self.ctx, self.klass, b.n)
        self.ctx.bind()
        _lib.check(_lib.lib().ssq_counter_pack_count(self.handle, ptr(b.ascii), int(b.ascii.numel()), ptr(b.offsets),
                                                    b.n, ptr(words), ptr(lens)))
        if check:
            _batch.raise_for_report(self.ctx.sync(), b)
        return _batch.ShortSeqArray(self.ctx, self.klass, words, lens)

    def insert(self, arr, check=True):
        """Count an already packed ShortSeqArray."""
        if arr.klass != self.klass:
            raise TypeError("array class does not match the counter class")
        self.ctx.bind()
        _lib.check(_lib.lib().ssq_counter_insert(self.handle, ptr(arr.words), ptr(arr.lens), len(arr)))
        if check:
            _batch.raise_for_report(self.ctx.sync())

    def merge(self, words, lens, counts):
        """Add weighted keys (tensors on this device): counts[i] occurrences of key i."""
        self.ctx.bind()
        _lib.check(_lib.lib().ssq_counter_merge(self.handle, ptr(words), ptr(lens), ptr(counts), int(lens.numel())))
        _batch.raise_for_report(self.ctx.sync())

    def clear(self):
        """Forget every key (the table keeps its size)."""
        self.ctx.bind()
        _lib.check(_lib.lib().ssq_counter_clear(self.handle))

    def merge_raw(self, words_ptr, lens_ptr, counts_ptr, n):
        """merge() on raw device pointers (the receive buffers of distributed.PeerExchange)."""
        self.ctx.bind()
        _lib.check(_lib.lib().ssq_counter_merge(self.handle, int(words_ptr), int(lens_ptr), int(counts_ptr), int(n)))
        _batch.raise_for_report(self.ctx.sync())

    def regions(self):
        """Number of table regions (ShortSeq64 counters; 0 otherwise)."""
        n = C.c_int64()
        _lib.check(_lib.lib().ssq_counter_regions(self.handle, C.byref(n)))
        return int(n.value)

    def export_region_bases(self, n_parts, dst_table):
        """Write, for every partition p, the offsets of its regions inside the block export_to sends to p to the device
        pointer dst_table[p] (int64 device tensor [n_parts]; the pointers may be peer memory)."""
        assert dst_table.dtype == torch.int64 and dst_table.numel() == n_parts and dst_table.is_contiguous()
        self.ctx.bind()
        _lib.check(_lib.lib().ssq_counter_export_region_bases(self.handle, int(n_parts), dst_table.data_ptr()))

    def merge_regions_raw(self, words_ptr, lens_ptr, counts_ptr, n, block_counts, block_regions, region_bases_ptr, rb_stride):
        """Region-aligned merge of hash-ordered blocks (see ssq_counter_merge_regions)."""
        bc = np.ascontiguousarray(block_counts, dtype=np.int64)
        br = np.ascontiguousarray(block_regions, dtype=np.int64)
        assert int(bc.sum()) == int(n) and bc.size == br.size
        self.ctx.bind()
        _lib.check(_lib.lib().ssq_counter_merge_regions(self.handle, int(words_ptr), int(lens_ptr), int(counts_ptr), bc.ctypes.data,
                                                       br.ctypes.data, int(bc.size), int(region_bases_ptr), int(rb_stride)))
        _batch.raise_for_report(self.ctx.sync())

    def export_counts(self, n_parts):
        """Tuples per hash partition (device int64 tensor) without exporting them (ShortSeq64 counters)."""
        parts = self.ctx.empty((n_parts,), torch.int64)
        self.ctx.bind()
        _lib.check(_lib.lib().ssq_counter_export_counts(self.handle, int(n_parts), ptr(parts)))
        return parts

    def export_to(self, n_parts, dst_table, first_part=0):
        """Write partition p's tuples to the device pointers dst_table[0][p] (words), dst_table[1][p] (lens),
        dst_table[2][p] (counts); dst_table is an int64 device tensor [3, n_parts].  The pointers may be peer memory;
        the kernel walks the partitions cyclically from first_part (stagger it across ranks)."""
        assert dst_table.dtype == torch.int64 and tuple(dst_table.shape) == (3, n_parts) and dst_table.is_contiguous()
        self.ctx.bind()
        base = dst_table.data_ptr()
        _lib.check(_lib.lib().ssq_counter_export_to(self.handle, int(n_parts), int(first_part), base, base + 8 * n_parts,
                                                   base + 16 * n_parts))

    def track_first_index(self, arr, base_index=0):
        """Record the first occurrence index of every key over this packed batch (dict order)."""
        self.ctx.bind()
        _lib.check(_lib.lib().ssq_counter_first_index(self.handle, ptr(arr.words), ptr(arr.lens), len(arr), int(base_index)))

    # -- queries ------------------------------------------------------------------------------
    def __len__(self):
        n = C.c_int64()
        self.ctx.bind()
        _lib.check(_lib.lib().ssq_counter_size(self.handle, C.byref(n)))
        return int(n.value)

    def capacity(self):
        n = C.c_int64()
        _lib.check(_lib.lib().ssq_counter_capacity(self.handle, C.byref(n)))
        return int(n.value)

    def lookup(self, arr):
        """Counts of the keys in arr (0 when absent) -> int64 tensor."""
        out = self.ctx.empty((len(arr),), torch.int64)
        self.ctx.bind()
        _lib.check(_lib.lib().ssq_counter_lookup(self.handle, ptr(arr.words), ptr(arr.lens), len(arr), ptr(out)))
        _batch.raise_for_report(self.ctx.sync())
        return out

    def export(self, n_parts=1, with_first_index=False):
        """All (key, len, count) tuples grouped by hash partition.

        -> (ShortSeqArray keys, counts int64 tensor, first_idx int64 tensor or None, part_counts int64 tensor).
        Partition p (the owner of a key in a p-way multi-GPU merge) is the slice
        [sum(part_counts[:p]), sum(part_counts[:p+1])).
        """
        n = len(self)
        ctx = self.ctx
        words = ctx.empty((n,) if self.klass == CLASS_64 else (n, 3), torch.int64)
        lens = ctx.empty((n,), torch.uint8)
        counts = ctx.empty((n,), torch.int64)
        first = ctx.empty((n,), torch.int64) if with_first_index else None
        parts = ctx.empty((n_parts,), torch.int64)
        ctx.bind()
        _lib.check(_lib.lib().ssq_counter_export(self.handle, int(n_parts), ptr(words), ptr(lens), ptr(counts), ptr(first),
                                                ptr(parts)))
        _batch.raise_for_report(ctx.sync())
        return _batch.ShortSeqArray(ctx, self.klass, words, lens), counts, first, parts


def count_reads(reads, device=None):
    """Count a list of bytes of mixed lengths on the GPU.

    -> one (klass, words, lens, counts, first_index) tuple of host arrays per length class present.  Reads of 0..96 nt
    are deduplicated.  ShortSeqVar-length reads (97..1024 nt) come back as one entry PER OCCURRENCE with count 1 and
    first_index = its list position: that is what the reference's counter does with them (it takes the dict hash from
    the 8 bytes after the object head, which is the heap pointer of a ShortSeqVar -- counter.pyx:44, short_seq_var.pxd:15
    -- so equal sequences never meet; SURVEY trap T3).
    """
    h_ascii, h_off = gather_reads(reads)
    lens = np.diff(h_off)
    errors = []
    too_long = np.nonzero(lens > 1024)[0]
    if too_long.size:
        errors.append((int(too_long[0]), Exception(MSG_TOO_LONG)))
    groups = []
    for k, idx, sub_ascii, sub_off in _batch.split_by_class(h_ascii, h_off):
        if k == CLASS_VAR:
            b = _batch.ReadBatch.make(sub_ascii, sub_off, device)
            arr, rep = _batch._pack_raw(b, CLASS_VAR)
            if rep.code != _lib.OK:
                try:
                    _batch.raise_for_report(rep, b, idx)
                except Exception as e:  # noqa: BLE001 -- re-raised below in list order
                    errors.append((int(idx[int(rep.first_bad_read)]), e))
                continue
            w, l, wo = arr.to_host()
            groups.append((CLASS_VAR, (w, wo), l, np.ones(idx.size, dtype=np.int64), idx))
            continue
        ctr = DeviceCounter(k, expected_unique=idx.size, device=device)
        b = _batch.ReadBatch.make(sub_ascii, sub_off, ctr.ctx.device)
        arr = ctr.pack_count(b, check=False)
        rep = ctr.ctx.sync()
        if rep.code != _lib.OK:
            try:
                _batch.raise_for_report(rep, b, idx)
            except Exception as e:  # noqa: BLE001 -- re-raised below in list order
                errors.append((int(idx[int(rep.first_bad_read)]), e))
            continue
        ctr.track_first_index(arr)
        keys, counts, first, _ = ctr.export(1, with_first_index=True)
        w, l, _ = keys.to_host()
        groups.append((k, w, l, counts.cpu().numpy(), idx[first.cpu().numpy()]))   # first = position in the original list
    if errors:
        raise min(errors, key=lambda t: t[0])[1]
    return groups


def _fill_in_order(dst, groups, add=False):
    """Insert the per-class results (klass, words, lens, counts, first_index arrays) into the dict `dst` in
    first-occurrence order.  Boxing and the insertion loop run in C when the helper module is built
    (hostext/_fastbox.c), else in bulk Python (dict.update over a zip).  ShortSeq64/192 keys are inserted under their
    hash (the first block); ShortSeqVar keys under their identity, one entry per occurrence, as in the reference."""
    from ._runtime import fastbox
    fb = fastbox()
    with _NoGC():              # millions of small objects are born here: the cyclic collector would rescan them over and over
        _fill_in_order_nogc(dst, groups, add, fb)


def _fill_in_order_nogc(dst, groups, add, fb):
    # Everything is put into first-occurrence order BEFORE the objects are created (one argsort + a few numpy takes per
    # class): the keys are then boxed in their final order and the dict is filled front to back -- objects and dict
    # entries are touched in allocation order instead of at random (1.26 M keys: fill 0.63 -> 0.36 s).
    objs, counts, firsts, hashes = [], [], [], []
    for klass, w, l, cnt, fi in groups:
        fi = np.asarray(fi, dtype=np.int64)
        cnt = np.asarray(cnt, dtype=np.int64)
        order = np.argsort(fi)              # first-occurrence indices are distinct within a class: no need for a stable sort
        fi, cnt, l = fi[order], cnt[order], np.asarray(l)[order]
        if klass == CLASS_VAR:
            words, wo = w
            boxed = [_box(CLASS_VAR, tuple(int(x) for x in words[wo[i]: wo[i + 1]]), int(l[j])) for j, i in enumerate(order.tolist())]
            hs = np.fromiter(map(id, boxed), dtype=np.int64, count=len(boxed))
        else:
            wv = np.ascontiguousarray(np.asarray(w)[order]).view(np.uint64)
            if fb is not None:
                boxed = fb.box_many(ShortSeq64 if klass == CLASS_64 else ShortSeq192, wv.tobytes(),
                                    np.ascontiguousarray(l, dtype=np.int64).tobytes(), 1 if klass == CLASS_64 else 3)
            else:
                boxed = _box_many(klass, wv.view(np.int64), l)
            hs = (wv if wv.ndim == 1 else wv[:, 0]).view(np.int64).copy()
            hs[hs == -1] = -2
        objs.append(boxed)
        hashes.append(hs)
        counts.append(cnt)
        firsts.append(fi)
    if not objs:
        return
    if len(objs) == 1:                      # one class (the usual case): already in order
        objs, counts, hashes = objs[0], counts[0], hashes[0]
        order = np.arange(len(objs), dtype=np.int64)
    else:                                   # interleave the classes by first occurrence
        objs = [o for part in objs for o in part]
        order = np.argsort(np.concatenate(firsts), kind="stable")
        counts, hashes = np.concatenate(counts), np.concatenate(hashes)
    if fb is not None:
        fb.fill_counts(dst, objs, counts.tobytes(), hashes.tobytes(), order.astype(np.int64).tobytes(), bool(add))
        return
    order, counts = order.tolist(), counts.tolist()
    if add:
        for j in order:
            dict.__setitem__(dst, objs[j], dict.get(dst, objs[j], 0) + counts[j])
    else:
        dict.update(dst, zip([objs[j] for j in order], [counts[j] for j in order]))


class ShortSeqCounter(dict):
    """dict of ShortSeq -> count, filled from a list of bytes (reference counter.pyx:10-54).

    Like the reference, only `type(source) is list` is consumed, elements must be bytes, and
    the first bad read (in list order) aborts construction with the reference's exception.
    Items iterate in first-occurrence order.
    """

    def __init__(self, source=None):
        super().__init__()
        if type(source) is list:
            self._count_py_bytes_list(source)

    def __setitem__(self, key, val):
        if type(key) not in (ShortSeq64, ShortSeq192, ShortSeqVar):
            raise TypeError(f"{self.__class__} does not support {type(key)} keys")
        dict.__setitem__(self, key, val)

    def _count_py_bytes_list(self, it):
        if not it:
            return
        _fill_in_order(self, count_reads(it), add=len(self) != 0)   # adding to a non-empty counter sums the counts

    @classmethod
    def from_batch(cls, source, offsets=None, klass=None, device=None):
        """Count a class-homogeneous batch given as ASCII buffer + offsets (no list of bytes needed)."""
        b = _batch.ReadBatch.make(source, offsets, device)
        if klass is None:
            klass = CLASS_64 if b.n == 0 else _batch.class_of_length(int(b.lengths_host()[0]))
        ctr = DeviceCounter(klass, expected_unique=b.n, device=b.ctx.device)
        arr = ctr.pack_count(b)
        ctr.track_first_index(arr)
        keys, counts, first, _ = ctr.export(1, with_first_index=True)
        w, l, _ = keys.to_host()
        cnt, fi = counts.cpu().numpy(), first.cpu().numpy()
        self = cls()
        _fill_in_order(self, [(klass, w, l, cnt, fi)])
        return self


def _fastq_reads_host(data, upto=None):
    """Host restatement of the reference's line selection (for error messages and the oracle tests): the reads of a
    FASTQ buffer, each minus its last byte."""
    buf = bytes(data)
    reads, pos, count = [], 0, 1
    while pos < len(buf):
        nl = buf.find(b"\n", pos)                      # getline: lines end at '\n' only
        end = len(buf) if nl < 0 else nl + 1
        if count % 2 == 0 and count % 4 != 0:
            reads.append(buf[pos:end - 1])
            if upto is not None and len(reads) > upto:
                break
        pos, count = end, count + 1
    return reads


def read_and_count_fastq(filename, device=None, chunk_bytes=0):
    """Count the sequence lines of a FASTQ file (reference counter.pyx:57-71, fast_read.pyx:3-20) on the GPU.

    Every 4k+2-th line is a read; like the reference's `_from_chars`, the last byte of the line (the newline) is
    dropped unconditionally (SURVEY trap T9).  The file is shipped to the device as raw text: newline scan, line
    selection, gather, pack and count all run there (ssq_host_fastq_count); only the distinct sequences come back.
    """
    data = np.fromfile(filename, dtype=np.uint8)
    ctx = context(device)
    self = ShortSeqCounter()
    if data.size == 0:
        return self
    # Size ONE counter from the first record (bytes per record -> reads in the file, class of its read): files are
    # almost always class-homogeneous, and creating / destroying a second large table costs more than counting the file
    # (cudaFree of a few hundred MB: ~0.1 s).  A counter whose bound turns out too small grows by itself; a read of the
    # other class makes the call report SSQ_ERR_CLASS, and the file is then counted again with both counters.
    head = data[: 1 << 16].tobytes()
    nl, pos_ = [], -1
    for _ in range(4):
        pos_ = head.find(b"\n", pos_ + 1)
        if pos_ < 0:
            break
        nl.append(pos_)
    rec_bytes = nl[3] + 1 if len(nl) == 4 else max(int(data.size), 1)
    first_len = (nl[1] - nl[0] - 1) if len(nl) >= 2 else 0
    bound = int(min(max(data.size / rec_bytes * 1.1 + 1024, 1 << 15), 1 << 27))
    primary = CLASS_64 if first_len <= 32 else CLASS_192
    n_reads, n_longer, first_longer = C.c_int64(), C.c_int64(), C.c_int64()
    rep = _lib.Report()
    h = ctx.bind()
    counters = {primary: DeviceCounter(primary, expected_unique=bound, device=ctx.device)}

    def run():
        c64, c192 = counters.get(CLASS_64), counters.get(CLASS_192)
        _lib.check(_lib.lib().ssq_host_fastq_count(h, None if c64 is None else c64.handle, None if c192 is None else c192.handle, data.ctypes.data,
                                                   int(data.size), int(chunk_bytes), 1, C.byref(n_reads), C.byref(n_longer),
                                                   C.byref(first_longer), C.byref(rep)))

    run()
    if rep.code == _lib.ERR_CLASS:                          # mixed classes: count again with a counter for each
        other = CLASS_192 if primary == CLASS_64 else CLASS_64
        counters[primary].clear()
        counters[other] = DeviceCounter(other, expected_unique=bound, device=ctx.device)
        run()
    c64 = counters.get(CLASS_64)
    c192 = counters.get(CLASS_192)
    bad = rep.first_bad_read if rep.code != _lib.OK else -1
    if n_longer.value and (bad < 0 or first_longer.value < bad):
        read = _fastq_reads_host(data, upto=first_longer.value)[first_longer.value]
        if len(read) > 1024:
            raise Exception(MSG_TOO_LONG)
        raise NotImplementedError("read_and_count_fastq: reads longer than 96 nt (ShortSeqVar) are not counted -- the reference "
                                  "does not deduplicate them either (each occurrence becomes its own key)")
    if rep.code == _lib.ERR_TABLE_FULL:
        raise _lib.LibraryError("counter table overflow")
    if bad >= 0:
        read = _fastq_reads_host(data, upto=bad)[bad]
        raise Exception(bad_base_message(read))
    groups = []
    for klass, ctr in ((CLASS_64, c64), (CLASS_192, c192)):
        if ctr is None or len(ctr) == 0:
            continue
        keys, counts, first, _ = ctr.export(1, with_first_index=True)
        w, l, _ = keys.to_host()
        groups.append((klass, w, l, counts.cpu().numpy(), first.cpu().numpy()))
    _fill_in_order(self, groups)
    return self
