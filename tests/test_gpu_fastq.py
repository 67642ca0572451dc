"""GPU: FASTQ ingest -- newline scan, line selection, gather, pack and count on the device --
against the reference's line rule (counter.pyx:57-71, fast_read.pyx:3-20) restated on the host, and against the
unmodified reference itself when oracle/_ref is built."""
import collections

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_fastq(rng, n, lo, hi, pool=None, unterminated=False, weird_quality=True):
    lines, reads = [], []
    for i in range(n):
        if pool is not None:
            seq = pool[int(rng.integers(0, len(pool)))]
        else:
            L = int(rng.integers(lo, hi + 1))
            seq = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=L).tobytes())
        reads.append(seq)
        qual = bytes(rng.integers(33, 74, size=len(seq), dtype=np.uint8).tobytes())
        if weird_quality and len(qual) >= 2 and i % 7 == 0:
            qual = b"@" + qual[1:-1] + b"\x0b"                 # '@' first, vertical tab right before the newline
        lines += [b"@read%d some header" % i, seq, b"+", qual]
    text = b"\n".join(lines) + (b"" if unterminated else b"\n")
    return text, reads


def expected_counts(reads):
    c = collections.OrderedDict()
    for r in reads:
        k = r.decode()
        c[k] = c.get(k, 0) + 1
    return c


def as_str_dict(counter):
    return collections.OrderedDict((str(k), v) for k, v in counter.items())


@pytest.mark.parametrize("chunk", [0, 4096, 70000])
def test_fastq_counts_and_order(sq, tmp_path, chunk):
    rng = np.random.default_rng(5)
    pool = [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(L)).tobytes())
            for L in rng.integers(0, 97, size=700)] + [b""]
    text, reads = make_fastq(rng, 20000, 0, 96, pool=pool)
    p = tmp_path / "a.fastq"
    p.write_bytes(text)
    got = sq.read_and_count_fastq(str(p), chunk_bytes=chunk)
    exp = expected_counts(reads)
    assert as_str_dict(got) == exp                              # same keys, counts
    assert list(as_str_dict(got)) == list(exp)                  # first-occurrence order across both classes
    assert type(got) is sq.ShortSeqCounter


def test_fastq_single_class_and_truncated_tail(sq, tmp_path):
    rng = np.random.default_rng(6)
    text, reads = make_fastq(rng, 5000, 22, 22, unterminated=True)
    # the file ends in the middle of a fifth record: header + unterminated sequence line (its last base is dropped, T9)
    text += b"\n@tail\nACGTACGTAC"
    p = tmp_path / "b.fastq"
    p.write_bytes(text)
    got = sq.read_and_count_fastq(str(p), chunk_bytes=8192)
    assert as_str_dict(got) == expected_counts(reads + [b"ACGTACGTA"])


def test_fastq_errors(sq, tmp_path):
    rng = np.random.default_rng(7)
    text, reads = make_fastq(rng, 3000, 15, 40)
    lines = text.split(b"\n")
    lines[4 * 1234 + 1] = b"ACGTNACGTACGTACGTACG"
    lines[4 * 2000 + 1] = b"ACGTNNNN"
    p = tmp_path / "c.fastq"
    p.write_bytes(b"\n".join(lines))
    with pytest.raises(Exception, match="Unsupported base character"):
        sq.read_and_count_fastq(str(p), chunk_bytes=16384)
    lines[4 * 100 + 1] = b"A" * 150                             # a ShortSeqVar-length read before the bad base
    p.write_bytes(b"\n".join(lines))
    with pytest.raises(NotImplementedError):
        sq.read_and_count_fastq(str(p))
    (tmp_path / "e.fastq").write_bytes(b"")
    assert len(sq.read_and_count_fastq(str(tmp_path / "e.fastq"))) == 0


def test_fastq_matches_reference(sq, tmp_path):
    from oracle import ref as R
    ref = R.load()
    if ref is None:
        pytest.skip("oracle/_ref (the built reference) is not available")
    rng = np.random.default_rng(8)
    pool = [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(L)).tobytes()) for L in rng.integers(1, 97, size=300)]
    text, _ = make_fastq(rng, 8000, 1, 96, pool=pool, weird_quality=False)
    p = tmp_path / "d.fastq"
    p.write_bytes(text)
    got = as_str_dict(sq.read_and_count_fastq(str(p), chunk_bytes=32768))
    exp = collections.OrderedDict((str(k), v) for k, v in ref.read_and_count_fastq(str(p)).items())
    assert got == exp and list(got) == list(exp)
