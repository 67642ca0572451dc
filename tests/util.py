"""Helpers shared by the tests."""
import numpy as np


def rand_reads(rng, n, lo, hi, alphabet=b"ACGT"):
    """n random reads with lengths uniform in [lo, hi] -> list of bytes."""
    lens = rng.integers(lo, hi + 1, size=n)
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    flat = alpha[rng.integers(0, len(alphabet), size=int(lens.sum()))].tobytes()
    out, pos = [], 0
    for L in lens:
        out.append(flat[pos:pos + L])
        pos += L
    return out


def concat(reads):
    lens = np.fromiter((len(r) for r in reads), dtype=np.int64, count=len(reads))
    offsets = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    return np.frombuffer(b"".join(reads), dtype=np.uint8).copy(), offsets


def counter_dict(words, lens, counts):
    """{(len, words tuple): count} from export arrays (words [n] or [n, W])."""
    words = np.asarray(words)
    out = {}
    for i in range(len(lens)):
        w = (int(words[i]),) if words.ndim == 1 else tuple(int(x) for x in words[i])
        key = (int(lens[i]), w)
        assert key not in out, "duplicate key in export"
        out[key] = int(counts[i])
    return out


def canon_counts(words, lens, counts):
    """Export arrays of a ShortSeq64 counter in a canonical order (by word, then length) -> (words, lens, counts);
    the vectorised stand-in for counter_dict when there are tens of millions of keys."""
    words = np.asarray(words, dtype=np.uint64)
    lens = np.asarray(lens).astype(np.int32)
    counts = np.asarray(counts).astype(np.int64)
    order = np.lexsort((lens, words))
    return words[order], lens[order], counts[order]
