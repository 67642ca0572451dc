"""GPU parity of the batched slicing / k-mer kernels and the tolerant alphabet against the oracle
and against the slices the unmodified reference produced (tests/golden/ref_vectors.json)."""
import json
import os

import numpy as np
import pytest

from tests.util import concat, rand_reads

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.json")
CLASS_RANGE = {0: (0, 32), 1: (33, 96), 2: (97, 1024)}


def _class_of(n):
    return 0 if n <= 32 else (1 if n <= 96 else 2)


def _expect_packed(oracle, seqs, klass):
    """Oracle packing of a list of (possibly short) sequences into the array layout of `klass`."""
    n = len(seqs)
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    if klass == 2:
        nb = (lens + 31) // 32
        off = np.zeros(n + 1, np.int64)
        np.cumsum(nb, out=off[1:])
        words = np.zeros(int(off[-1]), np.uint64)
        for i, s in enumerate(seqs):
            if s:
                words[off[i]: off[i + 1]] = oracle.pack_one(s)[1][: nb[i]]
        return words, lens, off
    W = 1 if klass == 0 else 3
    words = np.zeros((n, W), np.uint64)
    for i, s in enumerate(seqs):
        if s:
            w = oracle.pack_one(s)[1]
            words[i, : min(W, len(w))] = w[:W]
    return (words[:, 0] if W == 1 else words), lens, None


@pytest.mark.parametrize("klass", [0, 1, 2])
def test_slice_batch_scalar_bounds(sq, oracle, klass):
    rng = np.random.default_rng(40 + klass)
    lo, hi = CLASS_RANGE[klass]
    reads = rand_reads(rng, 3000, max(lo, 1), hi)
    arr = sq.pack_batch(reads, klass=klass)
    for start, stop in ((0, None), (0, 1), (3, 20), (5, 5), (7, 3), (10, 64), (31, 33), (0, 97), (40, 200), (100, 1024), (-5, None),
                        (-40, -3), (None, -1), (-2000, 10)):
        out = sq.slice_batch(arr, start, stop)
        want = [r[start:stop] for r in reads]
        width = max(len(x) for x in want)
        # the array class follows the largest possible slice (scalar, non-negative bounds: min(stop, class max) - start)
        ew, el, eo = _expect_packed(oracle, want, out.klass)
        w, l, wo = out.to_host()
        assert np.array_equal(l.astype(np.int64), el), (start, stop)
        assert np.array_equal(w, ew), (start, stop)
        if out.klass == 2:
            assert np.array_equal(wo, eo)
        assert width <= {0: 32, 1: 96, 2: 1024}[out.klass]
        # boxed elements take the class of their own length and compare equal to a fresh pack
        for i in (0, 17, 2999):
            o = out[i]
            assert type(o).__name__ == ("ShortSeq64", "ShortSeq192", "ShortSeqVar")[_class_of(len(want[i]))]
            assert o == sq.pack(want[i])


def test_slice_batch_per_read_bounds(sq, oracle):
    rng = np.random.default_rng(77)
    reads = rand_reads(rng, 5000, 97, 400)
    arr = sq.pack_batch(reads, klass=2)
    lens = np.array([len(r) for r in reads])
    starts = rng.integers(0, lens)
    widths = rng.integers(0, 90, size=len(reads))
    out = sq.slice_batch(arr, starts, starts + widths)
    want = [r[a:a + k] for r, a, k in zip(reads, starts.tolist(), widths.tolist())]
    assert out.klass == 1
    ew, el, _ = _expect_packed(oracle, want, 1)
    w, l, _ = out.to_host()
    assert np.array_equal(l.astype(np.int64), el) and np.array_equal(w, ew)
    # Hamming distance after slicing needs trimmed tails (reference unit_tests_main.py:402-435)
    a = sq.slice_batch(arr, 3, 35)
    b = sq.slice_batch(arr, 4, 36)
    d = sq.hamming_batch(a, b).cpu().numpy()
    assert np.array_equal(d, [sum(x != y for x, y in zip(r[3:35], r[4:36])) for r in reads])


def test_slices_of_the_reference(sq):
    """The 96 slices the unmodified reference produced: words, class and hash."""
    g = json.load(open(GOLDEN))["slices"]
    by_seq = {}
    for s in g:
        by_seq.setdefault(s["seq"], []).append(s)
    for seq, items in by_seq.items():
        arr = sq.pack_batch([seq.encode()] * len(items))
        starts = np.array([s["start"] for s in items], dtype=np.int64)
        stops = np.array([s["stop"] for s in items], dtype=np.int64)
        out = sq.slice_batch(arr, starts, stops)
        for i, s in enumerate(items):
            o = out[i]
            assert type(o).__name__ == s["type"]
            nb = len(s["words"])
            assert [hex(x) for x in o._packed[:nb]] == s["words"]
            assert hash(o) == s["hash"]


@pytest.mark.parametrize("klass,k,stride", [(0, 8, 1), (0, 21, 3), (1, 31, 1), (1, 32, 5), (2, 16, 7), (2, 32, 1)])
def test_kmers_batch(sq, oracle, klass, k, stride):
    rng = np.random.default_rng(500 + k)
    lo, hi = CLASS_RANGE[klass]
    reads = rand_reads(rng, 2000, max(lo, 1), min(hi, 300))
    arr = sq.pack_batch(reads, klass=klass)
    kmers, off = sq.kmers_batch(arr, k, stride)
    want, woff = [], [0]
    for r in reads:
        km = [r[j:j + k] for j in range(0, len(r) - k + 1, stride)] if len(r) >= k else []
        want += km
        woff.append(woff[-1] + len(km))
    assert np.array_equal(off.cpu().numpy(), woff)
    ew, el, _ = _expect_packed(oracle, want, 0)
    w, l, _ = kmers.to_host()
    assert np.array_equal(w, ew) and np.array_equal(l.astype(np.int64), el)
    # k-mer counting = counting the k-mer array
    ctr = sq.DeviceCounter(0, expected_unique=max(1024, len(want)))
    ctr.insert(kmers)
    from collections import Counter
    assert len(ctr) == len(Counter(want))


def test_tolerant_alphabet(sq, oracle):
    reads = [b"acgtACGTuU", b"uuuuUUUU" * 5, b"ACGT" * 30, b"gattaca", b"ACGN", b"acgx"]
    nb = sq.normalize_batch(reads[:4])
    upper = [r.upper().replace(b"U", b"T") for r in reads[:4]]
    assert bytes(nb.ascii.cpu().numpy()) == b"".join(upper)
    for klass, sub in ((0, [0, 3]), (1, [1]), (2, [2])):
        got = sq.pack_batch([bytes(nb.ascii.cpu().numpy()[nb.offsets[i]: nb.offsets[i + 1]]) for i in sub], klass=klass)
        buf, off = concat([upper[i] for i in sub])
        ow, ol, _ = oracle.pack_batch(klass, buf, off)
        w, l, _ = got.to_host()
        assert np.array_equal(w, ow) and np.array_equal(l.astype(np.int64), ol.astype(np.int64))
    # everything else is still rejected, and without the opt-in lower case / U are rejected as in the reference
    with pytest.raises(Exception, match="Unsupported base character"):
        sq.pack_batch(sq.normalize_batch(reads[4:5]))
    with pytest.raises(Exception, match="Unsupported base character"):
        sq.pack_batch(sq.normalize_batch(reads[5:6]))
    with pytest.raises(Exception, match="Unsupported base character"):
        sq.pack_batch(reads[:1])
