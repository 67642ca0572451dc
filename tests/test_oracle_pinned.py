"""CPU: pin the oracle (oracle/ssq_oracle.c) to the reference.

(a) SURVEY section 8c known-answer vectors, (b) golden fixtures generated from the unmodified
reference (tests/golden/ref_vectors.json, generator tests/golden/make_golden.py), and (c) the
live reference when oracle/_ref has been built in this container.
"""
import json
import os
import random

import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref as R

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.json")))
TYPE_OF_CLASS = {0: "ShortSeq64", 1: "ShortSeq192", 2: "ShortSeqVar"}


def test_kats():
    kats = [("A", [0x0]), ("C", [0x1]), ("T", [0x2]), ("G", [0x3]), ("ACGT", [0xb4]), ("ATGC", [0x78]),
            ("GATTACA", [0x4a3]), ("TGACTGACTGAC", [0x4e4e4e]), ("TGAGGTAGTAGGTTGTATAGTT", [0xac8baf2cbce]),
            ("GCGTAATAGGGGGTTTCGCTGTGGGGCGGCTAG", [0x27dffb9dabff20b7, 0x3, 0x0]), ("G" * 32, [0xffffffffffffffff]),
            ("ACGT" * 8, [0xb4b4b4b4b4b4b4b4]), ("A" * 32 + "C", [0, 1, 0]),
            ("TATTAGCGATTGACAGTTGTCCTGTAATAACGCCGGGTAAATTTGCCG", [0xd082e5bac4e8dca2, 0xd7a80bf5, 0x0]),
            ("TATTACCGATTGACAGTTGTCCTGTAATAACGGCGGGTAAATTTGCTG", [0xd082e5bac4e8d4a2, 0xe7a80bf7, 0x0]),
            ("ACGT" * 24, [0xb4b4b4b4b4b4b4b4] * 3), ("ACGT" * 24 + "T", [0xb4b4b4b4b4b4b4b4] * 3 + [0x2])]
    for s, words in kats:
        _, w = O.pack_one(s.encode())
        assert [int(x) for x in w] == words, s
    assert O.pyhash(0xffffffffffffffff) == -2 and O.pyhash(0xb4b4b4b4b4b4b4b4) == -5425512962855750476
    assert O.pyhash(0xd082e5bac4e8dca2) == -3421920176517882718


def test_golden_pack_hash_decode():
    for e in GOLDEN["pack"]:
        s = e["seq"].encode()
        k, w = O.pack_one(s)
        assert TYPE_OF_CLASS[k] == e["type"]
        assert [f"{int(x):#x}" for x in w][: len(e["words"])] == e["words"], e["seq"]
        assert O.pyhash(w[0]) == e["hash"]
        out, _ = O.decode_batch(w, np.array([len(s)]), stride=len(w), word_off=None)
        assert out.tobytes().decode() == e["str"] == e["seq"]


def test_golden_hamming():
    for e in GOLDEN["hamming"]:
        _, a = O.pack_one(e["a"].encode())
        _, b = O.pack_one(e["b"].encode())
        L = np.array([len(e["a"])])
        assert int(O.hamming_batch(a, b, L, L, len(a))[0]) == e["dist"]


def test_golden_counters():
    for case in GOLDEN["counters"]:
        reads = [r.encode() for r in case["reads"]]
        got = []
        for klass, stride in ((0, 1), (1, 3)):
            idx = [i for i, r in enumerate(reads) if O.lib().ssq_oracle_class(len(r)) == klass]
            if not idx:
                continue
            buf, off = O.concat([reads[i] for i in idx])
            w, l, _ = O.pack_batch(klass, buf, off)
            uw, ul, uc, fi = O.count(w, l, stride)
            for j in range(len(ul)):
                words = np.atleast_1d(uw[j])
                s, _ = O.decode_batch(words, np.array([ul[j]]), stride=stride)
                got.append((idx[int(fi[j])], s.tobytes().decode(), TYPE_OF_CLASS[klass], int(uc[j])))
        got.sort()
        assert [[s, t, c] for _, s, t, c in got] == case["items"]


def test_golden_rejects():
    for e in GOLDEN["rejects"]:
        with pytest.raises(O.OracleError) as ei:
            O.pack_one(e["seq"].encode())
        if "longer than" in e["message"]:
            assert ei.value.code == O.ERR_TOO_LONG
        else:
            assert ei.value.code == O.ERR_BAD_BASE
            assert e["message"] == "Unsupported base character: " + ei.value.bad_chars.decode()


def test_bloom_aliases_flagged_as_undefined():
    """SURVEY trap T1: 16 byte values pass the reference's bloom filter with undefined results."""
    alias = [c for c in range(256) if O.lib().ssq_oracle_is_base(c) and bytes([c]) not in (b"A", b"C", b"G", b"T")]
    assert len(alias) == 12 or len(alias) == 16 or len(alias) > 0
    for c in alias:
        with pytest.raises(O.OracleError) as ei:
            O.pack_one(bytes([c]))
        assert ei.value.code == O.ERR_UB
    for c in b"ACGT":
        assert O.lib().ssq_oracle_is_base(c)
    for c in b"NUacgtn*- ":
        assert not O.lib().ssq_oracle_is_base(c)


def test_batch_errors_and_classes():
    buf, off = O.concat([b"ACGT", b"ACNT", b"AC*T"])
    with pytest.raises(O.OracleError) as ei:
        O.pack_batch(0, buf, off)
    assert ei.value.first_bad == 1 and ei.value.bad_chars == b"N"
    buf, off = O.concat([b"ACGT", b"A" * 40])
    with pytest.raises(O.OracleError) as ei:
        O.pack_batch(0, buf, off)
    assert ei.value.code == O.ERR_CLASS and ei.value.first_bad == 1
    w, l, wo = O.pack_batch(2, *O.concat([b"A" * 97, b"C" * 1024, b"G" * 129]))
    assert list(wo) == [0, 4, 36, 41] and list(l) == [97, 1024, 129]


def test_synth_generator_is_deterministic_and_keyed():
    a1, o1 = O.synth_reads(0x5EED0001, 0, 2000, 50, 18, 30)
    a2, o2 = O.synth_reads(0x5EED0001, 1000, 1000, 50, 18, 30)
    assert np.array_equal(a1[o1[1000]:], a2) and np.array_equal(o1[1000:] - o1[1000], o2)
    reads = {a1[o1[i]:o1[i + 1]].tobytes() for i in range(2000)}
    assert len(reads) <= 50 and set(a1.tobytes()) <= set(b"ACGT")


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_live_reference_random():
    sq = R.load()
    rng = random.Random(5)
    reads = []
    for L in list(range(0, 140)) + [300, 1000, 1024]:
        s = "".join(rng.choice("ACGT") for _ in range(L))
        reads.append(s)
        o = sq.pack(s)
        k, w = O.pack_one(s.encode())
        rw = R.raw_words(o)
        assert [int(x) for x in w][: len(rw)] == rw and TYPE_OF_CLASS[k] == type(o).__name__
        assert O.pyhash(w[0]) == hash(o)
    pool = [r.encode() for r in reads[10:33]]
    lst = [rng.choice(pool) for _ in range(2000)]
    c = sq.ShortSeqCounter(lst)
    buf, off = O.concat(lst)
    w, l, _ = O.pack_batch(0, buf, off)
    uw, ul, uc, fi = O.count(w, l, 1)
    assert [(str(k), v) for k, v in c.items()] == [(lst[int(fi[j])].decode(), int(uc[j])) for j in range(len(ul))]
