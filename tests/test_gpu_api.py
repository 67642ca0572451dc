"""GPU: the reference's per-object Python API (sq.pack, ShortSeq*, ShortSeqCounter) as a drop-in.

Mirrors shortseq/tests/unit_tests_main.py of the reference.
"""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rand_seq(rng, n):
    return "".join(rng.choice("ACGT") for _ in range(n))


def test_empty_seq(sq):
    a, b = sq.pack(""), sq.pack(b"")
    assert a is b and str(a) == "" and len(a) == 0 and a == ""
    assert isinstance(a, sq.ShortSeq64)


def test_single_bases_and_types(sq):
    for c in "ACGT":
        s = sq.pack(c)
        assert isinstance(s, sq.ShortSeq64) and s == c and str(s) == c and len(s) == 1
        assert sq.pack(c.encode()) == c
    assert isinstance(sq.pack("A" * 32), sq.ShortSeq64)
    assert isinstance(sq.pack("A" * 33), sq.ShortSeq192)
    assert isinstance(sq.pack("A" * 96), sq.ShortSeq192)
    assert isinstance(sq.pack("A" * 97), sq.ShortSeqVar)
    assert isinstance(sq.pack("A" * 1024), sq.ShortSeqVar)
    with pytest.raises(TypeError, match="Cannot pack objects of type"):
        sq.pack(12)
    s = sq.pack("ACGT")
    assert sq.pack(s) is s


def test_kats_hash_eq_repr(sq):
    assert hash(sq.pack("ACGT")) == 180 and hash(sq.pack("ATGC")) == 120 and hash(sq.pack("GATTACA")) == 1187
    assert hash(sq.pack("G" * 32)) == -2
    assert hash(sq.pack("ACGT" * 8)) == -5425512962855750476
    assert hash(sq.pack("TATTAGCGATTGACAGTTGTCCTGTAATAACGCCGGGTAAATTTGCCG")) == -3421920176517882718
    assert hash(sq.pack("A")) == hash(sq.pack("AA")) == 0 and sq.pack("A") != sq.pack("AA")
    assert repr(sq.pack("ACGT")) == "<ShortSeq64 (4 nt): ACGT>"
    v = sq.pack("ACGT" * 30)
    assert repr(v) == f"<ShortSeqVar (120 nt): {('ACGT' * 30)[:75]} ... >"
    assert v.__sizeof__() == 32 + 8 * 4
    assert sq.pack("ACGT") != b"ACGT"      # reference quirk T6
    assert (sq.pack("ACGT") == 5) is False


def test_hamming_operator(sq):
    a = sq.pack("TATTAGCGATTGACAGTTGTCCTGTAATAACGCCGGGTAAATTTGCCG")
    b = sq.pack("TATTACCGATTGACAGTTGTCCTGTAATAACGGCGGGTAAATTTGCTG")
    assert a ^ b == 3
    assert sq.pack("A") ^ sq.pack("C") == 1 and sq.pack("A") ^ sq.pack("G") == 1 and sq.pack("ACGT") ^ sq.pack("ACGA") == 1
    with pytest.raises(Exception, match="equal length"):
        sq.pack("ACG") ^ sq.pack("ACGT")
    with pytest.raises(TypeError):
        sq.pack("ACG") ^ sq.pack("A" * 40)
    rng = random.Random(3)
    for L in (12, 31, 32, 33, 64, 95, 96, 97, 500, 1024):
        x, y = rand_seq(rng, L), rand_seq(rng, L)
        assert sq.pack(x) ^ sq.pack(y) == sum(p != q for p, q in zip(x, y))


def test_subscript_and_slices(sq):
    rng = random.Random(4)
    for L in (5, 32, 33, 70, 96, 97, 300):
        s = rand_seq(rng, L)
        p = sq.pack(s)
        for i in (0, 1, L // 2, L - 1, -1, -L):
            assert p[i] == s[i] and isinstance(p[i], sq.ShortSeq64)
        with pytest.raises(IndexError):
            p[L]
        with pytest.raises(IndexError):
            p[-L - 1]
        with pytest.raises(TypeError, match="Slice step not supported"):
            p[::2]
        with pytest.raises(TypeError, match="Invalid index type"):
            p["a"]
        assert p[3:3] is sq.empty
        for _ in range(40):
            a = rng.randrange(0, L); b = rng.randrange(a, L + 1)
            sl = p[a:b]
            assert str(sl) == s[a:b] and len(sl) == b - a
            n = b - a
            want = sq.ShortSeq64 if n <= 32 else sq.ShortSeq192 if n <= 96 else sq.ShortSeqVar
            assert isinstance(sl, want)
            if n:
                assert sl == sq.pack(s[a:b]) and hash(sl) == hash(sq.pack(s[a:b]))
    # Hamming after slicing needs trimmed tails (reference unit_tests_main.py:402-435)
    x, y = sq.pack("ACGT" * 10), sq.pack("ACGA" * 10)
    assert x[2:30] ^ y[2:30] == 7
    assert list(zip(sq.pack("ACG"), "ACG")) == [(sq.pack("A"), "A"), (sq.pack("C"), "C"), (sq.pack("G"), "G")]


def test_counter_readme_and_kats(sq):
    c = sq.ShortSeqCounter([b"ATGC"] * 10)
    assert c == {sq.pack("ATGC"): 10}
    c = sq.ShortSeqCounter([b"ACGT", b"TTTT", b"ACGT", b"GG", b"TTTT", b"ACGT"])
    assert [(str(k), v) for k, v in c.items()] == [("ACGT", 3), ("TTTT", 2), ("GG", 1)]
    c = sq.ShortSeqCounter([b"A", b"AA", b"AAA", b"A"])
    assert [(str(k), v) for k, v in c.items()] == [("A", 2), ("AA", 1), ("AAA", 1)]
    c = sq.ShortSeqCounter([b"", b"", b"A"])
    assert [(str(k), v) for k, v in c.items()] == [("", 2), ("A", 1)]
    assert sq.ShortSeqCounter((b"ACGT",)) == {}            # only lists are consumed (counter.pyx:14)
    with pytest.raises(TypeError, match="expected bytes, str found"):
        sq.ShortSeqCounter(["ACGT"])
    with pytest.raises(TypeError, match="does not support"):
        sq.ShortSeqCounter()["ACGT"] = 1
    with pytest.raises(Exception, match="Unsupported base character: N"):
        sq.ShortSeqCounter([b"ACGT", b"ACNT", b"AC*T"])
    # mixed 64 / 192 classes keep global first-occurrence order
    long = b"ACGT" * 12
    c = sq.ShortSeqCounter([long, b"ACGT", long, b"GG", b"ACGT"])
    assert [(str(k), v) for k, v in c.items()] == [(long.decode(), 2), ("ACGT", 2), ("GG", 1)]


def test_counter_from_batch_matches_collections_counter(sq):
    import collections
    rng = random.Random(9)
    pool = [rand_seq(rng, rng.randrange(15, 31)).encode() for _ in range(500)]
    reads = [rng.choice(pool) for _ in range(20_000)]
    c = sq.ShortSeqCounter(reads)
    ref = collections.Counter(reads)
    assert len(c) == len(ref)
    assert [str(k).encode() for k in c] == list(ref)            # same first-occurrence order
    assert all(c[sq.pack(k)] == v for k, v in list(ref.items())[:50])


def test_decode_many_and_hamming_many(sq):
    """One kernel per class instead of one launch per object; same answers as str() and ^ on the objects."""
    rng = np.random.default_rng(321)
    texts = ["".join(rng.choice(list("ACGT"), size=int(L))) for L in rng.integers(1, 300, size=400)]
    seqs = [sq.pack(t) for t in texts]
    assert sq.decode_many(seqs) == texts
    assert sq.decode_many([]) == []
    other = [sq.pack("".join(rng.choice(list("ACGT"), size=len(t)))) for t in texts]
    d = sq.hamming_many(seqs, other)
    assert d[:40] == [a ^ b for a, b in zip(seqs[:40], other[:40])]
    assert all(x == sum(c1 != c2 for c1, c2 in zip(t, str_o)) for x, t, str_o in zip(d, texts, sq.decode_many(other)))
    with pytest.raises(Exception, match="equal length"):
        sq.hamming_many([sq.pack("ACGT")], [sq.pack("ACG")])


def test_classify(sq):
    """ssq_classify: reads per container class from the offsets (reference short_seq.pyx:54-74)."""
    import torch
    from shortseq_b200 import _lib
    from shortseq_b200._runtime import context, ptr
    rng = np.random.default_rng(11)
    lens = rng.choice([0, 1, 31, 32, 33, 96, 97, 500, 1024, 1025, 4000], size=50_000)
    off = np.zeros(lens.size + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    ctx = context()
    d_off = torch.from_numpy(off).to(ctx.device)
    out = torch.empty(5, dtype=torch.int64, device=ctx.device)
    _lib.check(_lib.lib().ssq_classify(ctx.bind(), ptr(d_off), int(lens.size), ptr(out)))
    got = out.cpu().numpy()
    exp = [(lens <= 32).sum(), ((lens > 32) & (lens <= 96)).sum(), ((lens > 96) & (lens <= 1024)).sum(), (lens > 1024).sum(),
           int(np.nonzero(lens > 96)[0][0])]
    assert got.tolist() == [int(x) for x in exp]


def test_counter_shortseqvar_keys_like_the_reference(sq):
    """The reference's counter never deduplicates ShortSeqVar keys: it inserts them under the 8 bytes after the object
    head, a heap pointer (counter.pyx:44, short_seq_var.pxd:15; SURVEY trap T3).  Same here: one entry per occurrence,
    count 1, in list order, next to deduplicated ShortSeq64/192 keys."""
    a, b = b"ACGT" * 40, b"TTGCA" * 30
    reads = [a, b"ACGT", b, a, b"ACGT", b"G" * 50, a, b"G" * 50]
    c = sq.ShortSeqCounter(reads)
    assert [(str(k), v) for k, v in c.items()] == [(a.decode(), 1), ("ACGT", 2), (b.decode(), 1), (a.decode(), 1), ("G" * 50, 2),
                                                   (a.decode(), 1)]
    assert [type(k).__name__ for k in c] == ["ShortSeqVar", "ShortSeq64", "ShortSeqVar", "ShortSeqVar", "ShortSeq192", "ShortSeqVar"]
    from oracle import ref as R
    ref = R.load()
    if ref is not None:
        rc = ref.ShortSeqCounter(reads)
        assert [(str(k), v) for k, v in rc.items()] == [(str(k), v) for k, v in c.items()]
    with pytest.raises(Exception, match="longer than 1024"):
        sq.ShortSeqCounter([b"ACGT", b"A" * 1025])
    # adding a second list to a non-empty counter sums the counts of the deduplicated classes
    c2 = sq.ShortSeqCounter([b"ACGT", b"GG"])
    c2._count_py_bytes_list([b"GG", b"GG", b"T"])
    assert {str(k): v for k, v in c2.items()} == {"ACGT": 1, "GG": 3, "T": 1}


def test_single_object_entry_points_vs_oracle(sq, oracle):
    """ssq_pack_one / ssq_decode_one / ssq_hamming_one through the C ABI (ctypes): every length 1..1024 (step 7, plus the
    class boundaries) against the oracle's pack_one, decode round trip, Hamming against a mutated copy, bad bases."""
    import ctypes as C
    from shortseq_b200 import _lib
    from shortseq_b200._runtime import context
    lib, ctx = _lib.lib(), context()
    rng = random.Random(11)
    wa, wb = (C.c_uint64 * 32)(), (C.c_uint64 * 32)()
    klass, bad, dist = C.c_int32(), C.c_int32(), C.c_int32()
    text = C.create_string_buffer(1024)
    for n in sorted(set(list(range(1, 1025, 7)) + [31, 32, 33, 95, 96, 97, 1023, 1024])):
        s = rand_seq(rng, n).encode()
        assert lib.ssq_pack_one(ctx.handle, s, n, wa, C.byref(klass), C.byref(bad)) == 0
        oklass, ow = oracle.pack_one(s)
        nw = 1 if n <= 32 else (3 if n <= 96 else (n + 31) // 32)
        assert klass.value == oklass and bad.value == -1
        assert [int(x) for x in wa[:nw]] == [int(x) for x in ow[:nw]], n
        assert lib.ssq_decode_one(ctx.handle, wa, n, text) == 0 and text.raw[:n] == s
        # mutate k positions: distance k
        k = rng.randint(0, min(n, 9))
        t = bytearray(s)
        for p in rng.sample(range(n), k):
            t[p] = ord({"A": "C", "C": "G", "G": "T", "T": "A"}[chr(t[p])])
        assert lib.ssq_pack_one(ctx.handle, bytes(t), n, wb, C.byref(klass), None) == 0
        assert lib.ssq_hamming_one(ctx.handle, wa, wb, n, C.byref(dist)) == 0 and dist.value == k
    for n, pos in ((5, 0), (32, 31), (75, 40), (300, 299)):
        t = bytearray(rand_seq(rng, n).encode())
        t[pos] = ord("N")
        assert lib.ssq_pack_one(ctx.handle, bytes(t), n, wa, C.byref(klass), C.byref(bad)) == _lib.ERR_BAD_BASE
        assert bad.value == pos
        with pytest.raises(Exception, match="Unsupported base character"):
            sq.pack(bytes(t))
    assert lib.ssq_pack_one(ctx.handle, b"A", 0, wa, C.byref(klass), None) == _lib.ERR_ARG
    assert lib.ssq_pack_one(ctx.handle, b"A" * 1025, 1025, wa, C.byref(klass), None) == _lib.ERR_ARG
