"""CPU: host-side logic of the Python layer (no kernels): object protocol on boxed words, slicing,
error-message formatting, class splitting, partition hashing.  Packed words come from the oracle."""
import json
import os
import random

import numpy as np
import pytest

from oracle import oracle as O
from shortseq_b200 import hashing
from shortseq_b200._runtime import bad_base_message, gather_reads
from shortseq_b200.batch import class_of_length, split_by_class
from shortseq_b200.short_seq import ShortSeq64, ShortSeq192, ShortSeqVar, _box, empty

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.json")))
TYPES = {"ShortSeq64": ShortSeq64, "ShortSeq192": ShortSeq192, "ShortSeqVar": ShortSeqVar}


def box(seq: str):
    k, w = O.pack_one(seq.encode())
    return _box(k, [int(x) for x in w], len(seq))


def test_hash_len_eq_match_reference_golden():
    for e in GOLDEN["pack"]:
        o = box(e["seq"])
        assert type(o).__name__ == e["type"] and len(o) == e["len"] and hash(o) == e["hash"]
    a, aa = box("A"), box("AA")
    assert hash(a) == hash(aa) == 0 and a != aa and a == box("A") and not (a == 5)
    d = {box("ACGT"): 1}
    assert box("ACGT") in d and box("ACGA") not in d


def test_slices_match_reference_golden():
    for e in GOLDEN["slices"]:
        o = box(e["seq"])
        sl = o[e["start"]:e["stop"]]
        if e["stop"] == e["start"]:
            assert sl is empty
            continue
        assert type(sl) is TYPES[e["type"]]
        assert [f"{w:#x}" for w in sl._packed][: len(e["words"])] == e["words"]
        assert hash(sl) == e["hash"] and len(sl) == e["stop"] - e["start"]
        assert sl == box(e["seq"][e["start"]:e["stop"]])


def test_subscript_and_index_errors():
    rng = random.Random(1)
    for L in (1, 32, 33, 96, 97, 500):
        s = "".join(rng.choice("ACGT") for _ in range(L))
        o = box(s)
        for i in (0, L - 1, -1, -L, L // 2):
            assert o[i] == box(s[i]) and type(o[i]) is ShortSeq64
        with pytest.raises(IndexError, match="Sequence index out of range"):
            o[L]
        with pytest.raises(TypeError, match="Slice step not supported"):
            o[::2]
        with pytest.raises(TypeError, match="Invalid index type"):
            o[1.5]
    with pytest.raises(TypeError):
        ShortSeq64()


def test_bad_base_messages_match_reference_golden():
    for e in GOLDEN["rejects"]:
        if "longer than" in e["message"]:
            with pytest.raises(Exception, match="longer than 1024"):
                class_of_length(len(e["seq"]))
        else:
            assert bad_base_message(e["seq"].encode()) == e["message"]


def test_gather_and_split_by_class():
    reads = [b"ACGT", b"A" * 40, b"", b"C" * 200, b"GG", b"T" * 96, b"A" * 97]
    buf, off = gather_reads(reads)
    assert buf.tobytes() == b"".join(reads) and list(np.diff(off)) == [len(r) for r in reads]
    parts = {k: (idx, a, o) for k, idx, a, o in split_by_class(buf, off)}
    assert list(parts[0][0]) == [0, 2, 4] and list(parts[1][0]) == [1, 5] and list(parts[2][0]) == [3, 6]
    for k, (idx, a, o) in parts.items():
        assert [a[o[j]:o[j + 1]].tobytes() for j in range(len(idx))] == [reads[i] for i in idx]
    with pytest.raises(TypeError, match="expected bytes, str found"):
        gather_reads([b"A", "C"])


def test_owner_rank_partitions_evenly_and_by_full_key():
    rng = np.random.default_rng(0)
    w = rng.integers(0, 2**63, size=20000, dtype=np.int64).astype(np.uint64)
    l = np.full(20000, 32)
    for world in (1, 2, 4, 8):
        r = hashing.owner_rank(w, l, 0, world)
        assert r.min() >= 0 and r.max() < world
        assert np.bincount(r, minlength=world).min() > 20000 / world * 0.9
    w3 = np.stack([w, w[::-1], w ^ np.uint64(5)], axis=1)
    r1 = hashing.owner_rank(w3, np.full(20000, 75), 1, 8)
    r2 = hashing.owner_rank(w3, np.full(20000, 76), 1, 8)
    assert (r1 != r2).mean() > 0.5            # length is part of the key
    assert int(hashing.mix64(np.uint64(0))) == int(O.lib().ssq_oracle_mix64(0))
    assert int(hashing.mix64(np.uint64(12345))) == int(O.lib().ssq_oracle_mix64(12345))


def test_fastq_line_rule_matches_reference(tmp_path):
    """The host restatement of the reference's FASTQ line rule (shortseq_b200.counter._fastq_reads_host: lines 2 mod 4,
    last byte dropped, lines end at '\\n' only) -- the expectation the GPU ingest tests compare against -- agrees with the
    unmodified reference's read_and_count_fastq (counter.pyx:57-71) when oracle/_ref is built."""
    import collections
    from oracle import ref as R
    from shortseq_b200.counter import _fastq_reads_host
    ref = R.load()
    if ref is None:
        pytest.skip("oracle/_ref (the built reference) is not available")
    rng = np.random.default_rng(99)
    recs = []
    for i in range(3000):
        L = int(rng.integers(1, 97))
        seq = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=L).tobytes())
        qual = bytes(rng.integers(33, 74, size=L, dtype=np.uint8).tobytes())
        if i % 5 == 0:
            qual = b"@" + qual[1:]
        recs += [b"@r%d" % i, seq, b"+", qual]
    for tail in (b"\n", b"", b"\n@last\nACGTAC"):            # terminated, unterminated, truncated record (T9)
        text = b"\n".join(recs) + tail
        p = tmp_path / "x.fastq"
        p.write_bytes(text)
        mine = collections.OrderedDict()
        for r in _fastq_reads_host(np.frombuffer(text, dtype=np.uint8)):
            mine[r.decode()] = mine.get(r.decode(), 0) + 1
        theirs = collections.OrderedDict((str(k), v) for k, v in ref.read_and_count_fastq(str(p)).items())
        assert mine == theirs and list(mine) == list(theirs)


def test_streamed_exchange_key_conversion():
    """The host statement of the streamed exchange's key rewrite (ssq_counter.cu, count_regions2_kernel<., true>): a key
    read from a slot of the sender's table, together with its region index, becomes the owner table's key; the owner's
    region and slot offset are bits of that key, and the owner rank is the top log2(P) bits of the unrotated hash."""
    rng = np.random.default_rng(5)
    n = 20000
    lens = rng.integers(1, 33, size=n).astype(np.uint8)
    words = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) >> (np.uint64(64) - 2 * lens.astype(np.uint64))
    for world, log2_cap_local, log2_cap_owner in ((2, 28, 27), (8, 28, 25), (4, 22, 20), (1, 22, 22), (2, 28, 26)):
        rot_o = world.bit_length() - 1
        lr = 12
        log2_regions = log2_cap_local - lr
        h2, key = hashing.table_key64(words, lens, 0)
        region = (h2 >> np.uint64(64 - log2_cap_local)) >> np.uint64(lr)            # home region = the region a key is found in
        sent = hashing.streamed_key64(key, region, log2_regions, 0, rot_o)
        h2_o, key_o = hashing.table_key64(words, lens, rot_o)
        assert np.array_equal(sent, key_o)
        # the owner of the region the key sat in is the key's owner rank
        owner = region >> np.uint64(log2_regions - rot_o) if rot_o else np.zeros(n, np.uint64)
        assert np.array_equal(owner.astype(np.int64), hashing.owner_rank(words, lens, 0, world))
        # merge_regions_kernel<true>: region and offset in the owner table from the key alone (minus the six top bits)
        off_shift = 64 - log2_cap_owner
        slot_low = (sent & np.uint64((1 << 58) - 1)) >> np.uint64(off_shift)
        slot = h2_o >> np.uint64(off_shift)
        assert np.array_equal(slot_low & np.uint64((1 << lr) - 1), slot & np.uint64((1 << lr) - 1))
        assert np.array_equal(slot_low >> np.uint64(lr), (slot >> np.uint64(lr)) & np.uint64((1 << (log2_cap_owner - lr - 6)) - 1))
        # the sender's region in_owner index refines the owner's region index (the grids nest)
        in_owner = region & np.uint64((1 << (log2_regions - rot_o)) - 1)
        ratio = (log2_regions - rot_o) - (log2_cap_owner - lr)
        assert ratio >= 0 and np.array_equal(in_owner >> np.uint64(ratio), slot >> np.uint64(lr))
