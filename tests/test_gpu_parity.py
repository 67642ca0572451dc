"""GPU parity: every CUDA path through the C ABI against the CPU oracle, bit-exact."""
import numpy as np
import pytest
import torch

from tests.util import concat, counter_dict, rand_reads

pytestmark = pytest.mark.gpu

CLASS_RANGE = {0: (0, 32), 1: (33, 96), 2: (97, 1024)}


def _pack_and_compare(sq, oracle, reads, klass):
    buf, off = concat(reads)
    arr = sq.pack_batch(buf, off, klass=klass)
    w, l, wo = arr.to_host()
    ow, ol, owo = oracle.pack_batch(klass, buf, off)
    assert np.array_equal(l.astype(np.int64), ol.astype(np.int64))
    assert np.array_equal(w, ow)
    if klass == 2:
        assert np.array_equal(wo, owo)
    return arr


@pytest.mark.parametrize("klass", [0, 1, 2])
def test_pack_every_length(sq, oracle, klass):
    """reference tests: test_length_range / test_min_length / test_max_length (unit_tests_main.py:91-118,249-287)."""
    rng = np.random.default_rng(10 + klass)
    lo, hi = CLASS_RANGE[klass]
    reads = []
    for L in range(lo, hi + 1):
        reads += rand_reads(rng, 2 if klass == 2 else 5, L, L)
    rng.shuffle(reads)
    _pack_and_compare(sq, oracle, reads, klass)


@pytest.mark.parametrize("klass,n", [(0, 100_000), (1, 50_000), (2, 3_000)])
def test_pack_random_batches(sq, oracle, klass, n):
    rng = np.random.default_rng(20 + klass)
    lo, hi = CLASS_RANGE[klass]
    _pack_and_compare(sq, oracle, rand_reads(rng, n, lo, hi), klass)


@pytest.mark.parametrize("klass,L", [(0, 32), (0, 22), (0, 1), (1, 75), (1, 96), (2, 150), (2, 1024)])
def test_pack_fixed_length(sq, oracle, klass, L):
    rng = np.random.default_rng(30 + L)
    _pack_and_compare(sq, oracle, rand_reads(rng, 4099, L, L), klass)


def test_pack_homopolymers_and_kats(sq, oracle):
    reads = [b"A" * 32, b"C" * 32, b"G" * 32, b"T" * 32, b"ACGT", b"ATGC", b"GATTACA", b"TGACTGACTGAC",
             b"TGAGGTAGTAGGTTGTATAGTT", b"ACGT" * 8, b"", b"A", b"AA"]
    arr = _pack_and_compare(sq, oracle, reads, 0)
    w, _, _ = arr.to_host()
    assert [int(x) for x in w[4:9]] == [0xb4, 0x78, 0x4a3, 0x4e4e4e, 0xac8baf2cbce]
    assert int(w[2]) == 0xFFFFFFFFFFFFFFFF and int(w[9]) == 0xb4b4b4b4b4b4b4b4
    arr = _pack_and_compare(sq, oracle, [b"GCGTAATAGGGGGTTTCGCTGTGGGGCGGCTAG", b"A" * 32 + b"C", b"ACGT" * 24], 1)
    w, _, _ = arr.to_host()
    assert [int(x) for x in w[0]] == [0x27dffb9dabff20b7, 0x3, 0x0]
    assert [int(x) for x in w[1]] == [0, 1, 0]
    _pack_and_compare(sq, oracle, [b"ACGT" * 24 + b"T"], 2)


def test_pack_empty_batch(sq):
    arr = sq.pack_batch(np.zeros(0, np.uint8), np.zeros(1, np.int64), klass=0)
    assert len(arr) == 0


def test_pack_unaligned_device_buffer(sq, oracle):
    """The ASCII pointer handed to the C ABI need not be 16-byte aligned."""
    rng = np.random.default_rng(5)
    reads = rand_reads(rng, 3000, 0, 32)
    buf, off = concat(reads)
    for shift in (1, 3, 8, 15):
        dev = torch.zeros(len(buf) + 32, dtype=torch.uint8, device="cuda")
        dev[shift:shift + len(buf)] = torch.from_numpy(buf).cuda()
        arr = sq.pack_batch(dev[shift:shift + len(buf)], torch.from_numpy(off).cuda(), klass=0)
        ow, ol, _ = oracle.pack_batch(0, buf, off)
        w, l, _ = arr.to_host()
        assert np.array_equal(w, ow) and np.array_equal(l, ol)


def test_every_byte_value_validation(sq):
    """Exact {A,C,G,T} validation for all 256 byte values (the reference's bloom lets 16 aliases through)."""
    for c in range(256):
        for L, klass in ((1, 0), (20, 0), (40, 1), (130, 2)):
            read = bytearray(b"ACGT" * 40)[:L]
            read[L // 2] = c
            ok = bytes([c]) in (b"A", b"C", b"G", b"T")
            if ok:
                sq.pack_batch([bytes(read)], klass=klass)
            else:
                with pytest.raises(Exception, match="Unsupported base character"):
                    sq.pack_batch([bytes(read)], klass=klass)


def test_bad_base_reports_lowest_read_and_reference_message(sq, oracle):
    """reference: test_incompatible_seq_chars (unit_tests_main.py:63-69,504-515) + SURVEY KAT-20."""
    rng = np.random.default_rng(7)
    reads = rand_reads(rng, 5000, 10, 32)
    for bad_at in (4999, 2500, 17):
        r = bytearray(reads[bad_at]); r[len(r) // 2] = ord("N"); reads[bad_at] = bytes(r)
        with pytest.raises(Exception, match="Unsupported base character: N"):
            sq.pack_batch(reads, klass=0)
        from shortseq_b200 import _lib
        from shortseq_b200.batch import ReadBatch, _pack_raw
        _, rep = _pack_raw(ReadBatch.make(reads), 0)
        assert rep.code == _lib.ERR_BAD_BASE and rep.first_bad_read == bad_at
    # message forms of the reference
    for read, msg in [(b"N", "N"), (b"*", r"\*"), (b"U", "U"), (b"a", "a"), (b"acgt", "t"),
                      (b"A" * 32 + b"N", "N"), (b"A" * 8 + b"N" + b"A" * 25, "NAAAAAAA")]:
        with pytest.raises(Exception, match="Unsupported base character: " + msg):
            sq.pack(read)
        with pytest.raises(oracle.OracleError):
            oracle.pack_one(read)


def test_last_char_bad_every_var_length(sq):
    """reference unit_tests_main.py:504-515: last char bad for every L in 97..1023."""
    reads = [b"A" * (L - 1) + b"N" for L in range(97, 1024)]
    from shortseq_b200 import _lib
    from shortseq_b200.batch import ReadBatch, _pack_raw
    for i in (0, 500, len(reads) - 1):
        good = [b"A" * len(r) for r in reads]
        good[i] = reads[i]
        _, rep = _pack_raw(ReadBatch.make(good), 2)
        assert rep.code == _lib.ERR_BAD_BASE and rep.first_bad_read == i


def test_var_bad_reads_spread_over_tiles(sq):
    """ShortSeqVar pack finds the failing reads from a shared-memory list of invalid chunks: bad bytes at the first / last
    byte of a read (they share 16-byte chunks and 32-read tiles with their neighbours), many bad reads at once, and more
    invalid chunks in one tile than the list holds (the kernel then re-reads the tile's reads): the report is always the
    LOWEST failing read."""
    from shortseq_b200 import _lib
    from shortseq_b200.batch import ReadBatch, _pack_raw
    rng = np.random.default_rng(23)
    reads = rand_reads(rng, 4000, 97, 700)

    def first_bad(rs):
        _, rep = _pack_raw(ReadBatch.make(rs), 2)
        assert rep.code == _lib.ERR_BAD_BASE
        return int(rep.first_bad_read)

    for positions in ([(3999, -1)], [(3100, 0)], [(31, -1), (32, 0)], [(2000, 5), (64, 0), (63, -1), (3000, 7)],
                      [(i, int(rng.integers(0, 97))) for i in range(500, 4000, 37)]):
        rs = list(reads)
        for i, at in positions:
            r = bytearray(rs[i]); r[at] = ord("N"); rs[i] = bytes(r)
        assert first_bad(rs) == min(i for i, _ in positions)
    rs = list(reads)                                   # every read of tiles 40..41 bad in many places: list overflow
    for i in range(40 * 32, 42 * 32):
        r = bytearray(rs[i]); r[::9] = b"N" * len(r[::9]); rs[i] = bytes(r)
    assert first_bad(rs) == 40 * 32
    rs[7] = rs[7][:50] + b"n" + rs[7][51:]
    assert first_bad(rs) == 7


def test_length_errors(sq):
    with pytest.raises(Exception, match="longer than 1024 bases"):
        sq.pack(b"A" * 1025)
    with pytest.raises(sq.ShortSeqClassError):
        sq.pack_batch([b"A" * 10, b"A" * 40], klass=0)
    with pytest.raises(sq.ShortSeqClassError):
        sq.pack_batch([b"A" * 40, b"A" * 10], klass=1)
    with pytest.raises(Exception, match="longer than 1024 bases"):
        sq.pack_batch([b"A" * 100, b"A" * 2000], klass=2)


@pytest.mark.parametrize("klass,n", [(0, 50_000), (1, 20_000), (2, 2_000)])
def test_decode_round_trip(sq, oracle, klass, n):
    """reference: test_length_range round trips (str(pack(s)) == s)."""
    rng = np.random.default_rng(40 + klass)
    lo, hi = CLASS_RANGE[klass]
    reads = rand_reads(rng, n, lo, hi)
    buf, off = concat(reads)
    arr = sq.pack_batch(buf, off, klass=klass)
    out, out_off = arr.decode()
    assert np.array_equal(out_off.cpu().numpy(), off)
    assert np.array_equal(out.cpu().numpy(), buf)
    w, l, wo = arr.to_host()
    oa, _ = oracle.decode_batch(w, l, stride=3 if klass == 1 else 1, word_off=wo)
    assert np.array_equal(oa, buf)


@pytest.mark.parametrize("klass,n", [(0, 50_000), (1, 20_000), (2, 2_000)])
def test_hamming_pairs(sq, oracle, klass, n):
    """reference: test_hamming_distance (unit_tests_main.py:159-166,456-463)."""
    rng = np.random.default_rng(50 + klass)
    lo, hi = CLASS_RANGE[klass]
    a = rand_reads(rng, n, lo, hi)
    b = []
    for r in a:  # mutate a few positions
        m = bytearray(r)
        for _ in range(rng.integers(0, 6)):
            if m:
                m[rng.integers(0, len(m))] = b"ACGT"[rng.integers(0, 4)]
        b.append(bytes(m))
    A, B = sq.pack_batch(a, klass=klass), sq.pack_batch(b, klass=klass)
    d = sq.hamming_batch(A, B).cpu().numpy().astype(np.int64)
    expect = np.array([sum(x != y for x, y in zip(ra, rb)) for ra, rb in zip(a, b)])
    assert np.array_equal(d, expect)
    if klass != 2:
        aw, al, _ = A.to_host(); bw, bl, _ = B.to_host()
        od = oracle.hamming_batch(aw, bw, al, bl, 3 if klass == 1 else 1)
        assert np.array_equal(d, od)


def test_hamming_length_mismatch(sq):
    A, B = sq.pack_batch([b"ACGT", b"ACG"], klass=0), sq.pack_batch([b"ACGT", b"ACGT"], klass=0)
    with pytest.raises(Exception, match="equal length"):
        sq.hamming_batch(A, B)


def test_hamming_refset(sq):
    rng = np.random.default_rng(60)
    # 12 / 16 nt: the 32-bit UMI loop; 20 / 32 nt: one 64-bit word; 96 nt: three words.  1500 references = one uniform
    # chunk + one chunk with a second length in it; ties in the distance are common (first minimal index wins)
    for klass, L in ((0, 12), (0, 16), (0, 20), (0, 32), (1, 96)):
        refs = rand_reads(rng, 1500, L, L) + rand_reads(rng, 10, L - 1, L - 1)
        q = rand_reads(rng, 3000, L, L)
        Q, R = sq.pack_batch(q, klass=klass), sq.pack_batch(refs, klass=klass)
        md, am, within = sq.hamming_refset(Q, R, thresh=L // 2)
        qa = np.array([list(x) for x in q], dtype=np.uint8)
        ra = np.array([list(x) for x in refs[:1500]], dtype=np.uint8)
        dist = (qa[:, None, :] != ra[None, :, :]).sum(-1)
        assert np.array_equal(md.cpu().numpy(), dist.min(1))
        assert np.array_equal(am.cpu().numpy(), dist.argmin(1))
        assert np.array_equal(within.cpu().numpy(), (dist <= L // 2).sum(1))


@pytest.mark.parametrize("klass,n,u,lo,hi", [(0, 200_000, 5_000, 15, 32), (0, 100_000, 100_000, 32, 32),
                                            (0, 50_000, 40, 0, 3), (1, 100_000, 3_000, 33, 96),
                                            (1, 50_000, 50_000, 75, 75)])
def test_counter_vs_oracle(sq, oracle, klass, n, u, lo, hi):
    """reference: ShortSeqCounter (counter.pyx:41-54); key = (length, words)."""
    rng = np.random.default_rng(70 + klass + u)
    pool = rand_reads(rng, u, lo, hi)
    reads = [pool[i] for i in rng.integers(0, u, size=n)]
    buf, off = concat(reads)
    ctr = sq.DeviceCounter(klass, expected_unique=0)           # starts at the minimum size: exercises growth
    arr = ctr.pack_count(buf, off)
    ctr.track_first_index(arr)
    keys, counts, first, parts = ctr.export(1, with_first_index=True)
    kw, kl, _ = keys.to_host()
    ow, ol, _ = oracle.pack_batch(klass, buf, off)
    uw, ul, uc, ufi = oracle.count(ow, ol, 3 if klass == 1 else 1)
    assert len(ctr) == len(ul) == int(parts.sum())
    got = counter_dict(kw, kl, counts.cpu().numpy())
    exp = counter_dict(uw, ul, uc)
    assert got == exp
    gfirst = counter_dict(kw, kl, first.cpu().numpy())
    efirst = counter_dict(uw, ul, ufi)
    assert gfirst == efirst
    # lookup
    probe = sq.pack_batch(buf, off, klass=klass)
    c = ctr.lookup(probe).cpu().numpy()
    assert all(c[i] == exp[(int(ol[i]), (int(ow[i]),) if klass == 0 else tuple(int(x) for x in ow[i]))] for i in range(0, n, 997))


def test_counter_incremental_and_insert_packed(sq, oracle):
    rng = np.random.default_rng(81)
    pool = rand_reads(rng, 2000, 18, 30)
    ctr = sq.DeviceCounter(0, expected_unique=100)
    all_reads = []
    for step in range(4):
        reads = [pool[i] for i in rng.integers(0, 2000, size=30_000)]
        all_reads += reads
        if step % 2:
            ctr.insert(sq.pack_batch(reads, klass=0))
        else:
            ctr.pack_count(reads)
    buf, off = concat(all_reads)
    ow, ol, _ = oracle.pack_batch(0, buf, off)
    uw, ul, uc, _ = oracle.count(ow, ol, 1)
    keys, counts, _, _ = ctr.export(1)
    kw, kl, _ = keys.to_host()
    assert counter_dict(kw, kl, counts.cpu().numpy()) == counter_dict(uw, ul, uc)


@pytest.mark.parametrize("klass", [0, 1])
def test_counter_export_partitions_and_merge(sq, oracle, klass):
    """Hash-partitioned export + weighted merge = the single-GPU form of the multi-GPU all-to-all."""
    rng = np.random.default_rng(90 + klass)
    lo, hi = (20, 32) if klass == 0 else (40, 90)
    pool = rand_reads(rng, 20_000, lo, hi)
    shards = [[pool[i] for i in rng.integers(0, 20_000, size=60_000)] for _ in range(4)]
    P = 4
    owners = [sq.DeviceCounter(klass, expected_unique=8000, hash_rot=2) for _ in range(P)]
    for shard in shards:
        local = sq.DeviceCounter(klass, expected_unique=30_000)
        local.pack_count(shard)
        keys, counts, _, parts = local.export(P)
        parts = parts.cpu().numpy()
        assert parts.sum() == len(local)
        start = 0
        for p in range(P):
            sl = slice(start, start + int(parts[p]))
            owners[p].merge(keys.words[sl].contiguous(), keys.lens[sl].contiguous(), counts[sl].contiguous())
            start += int(parts[p])
    merged = {}
    for o in owners:
        keys, counts, _, _ = o.export(1)
        kw, kl, _ = keys.to_host()
        d = counter_dict(kw, kl, counts.cpu().numpy())
        assert not (set(d) & set(merged)), "a key landed on two owners"
        merged.update(d)
    buf, off = concat([r for s in shards for r in s])
    ow, ol, _ = oracle.pack_batch(klass, buf, off)
    uw, ul, uc, _ = oracle.count(ow, ol, 3 if klass == 1 else 1)
    assert merged == counter_dict(uw, ul, uc)


def test_counter_skewed_hot_keys(sq, oracle):
    """A few keys carry most of the reads (small-RNA like): contended atomics must stay exact."""
    rng = np.random.default_rng(99)
    pool = rand_reads(rng, 1000, 20, 24)
    idx = np.minimum((rng.pareto(0.7, size=300_000)).astype(np.int64), 999)
    reads = [pool[i] for i in idx]
    ctr = sq.DeviceCounter(0, expected_unique=2000)
    ctr.pack_count(reads)
    keys, counts, _, _ = ctr.export(1)
    kw, kl, _ = keys.to_host()
    buf, off = concat(reads)
    ow, ol, _ = oracle.pack_batch(0, buf, off)
    uw, ul, uc, _ = oracle.count(ow, ol, 1)
    assert counter_dict(kw, kl, counts.cpu().numpy()) == counter_dict(uw, ul, uc)


def test_synth_reads_match_oracle(sq, oracle):
    for lo, hi, n, u in ((32, 32, 10_000, 300), (22, 22, 5000, 100), (33, 96, 4000, 500), (97, 1024, 300, 50)):
        b = sq.synth_reads(n, u, lo, hi, seed=0x5EED0001, first_read=123)
        oa, oo = oracle.synth_reads(0x5EED0001, 123, n, u, lo, hi)
        assert np.array_equal(b.offsets.cpu().numpy(), oo)
        assert np.array_equal(b.ascii.cpu().numpy(), oa)


def test_host_pipeline(sq, oracle):
    """ssq_host_pack_count: host buffers in, packed words + counts out, chunked."""
    import ctypes as C
    from shortseq_b200 import _lib
    rng = np.random.default_rng(111)
    pool = rand_reads(rng, 3000, 10, 32)
    reads = [pool[i] for i in rng.integers(0, 3000, size=100_000)]
    buf, off = concat(reads)
    ctr = sq.DeviceCounter(0, expected_unique=4000)
    words = np.zeros(len(reads), np.uint64)
    lens = np.zeros(len(reads), np.uint8)
    rep = _lib.Report()
    _lib.check(_lib.lib().ssq_host_pack_count(ctr.ctx.bind(), ctr.handle, buf.ctypes.data, off.ctypes.data, len(reads),
                                             words.ctypes.data, lens.ctypes.data, 7777, C.byref(rep)))
    assert rep.code == 0
    ow, ol, _ = oracle.pack_batch(0, buf, off)
    assert np.array_equal(words, ow) and np.array_equal(lens, ol)
    uw, ul, uc, _ = oracle.count(ow, ol, 1)
    keys, counts, _, _ = ctr.export(1)
    kw, kl, _ = keys.to_host()
    assert counter_dict(kw, kl, counts.cpu().numpy()) == counter_dict(uw, ul, uc)


def _oracle_counts_of_batch(oracle, b, klass):
    buf, off = b.ascii.cpu().numpy(), b.offsets.cpu().numpy()
    ow, ol, _ = oracle.pack_batch(klass, buf, off)
    uw, ul, uc, _ = oracle.count(ow, ol, 3 if klass == 1 else 1)
    return (ow, ol), counter_dict(uw, ul, uc)


@pytest.mark.parametrize("fused", [True, False])
def test_counter_deferred_partition_path(sq, oracle, fused):
    """Tables too large for L2 take the two-phase path (scatter to 256 hash partitions, then insert
    partition by partition); results must be identical to the direct path and to the oracle."""
    n, u = 3_000_000, 1_500_000
    b = sq.synth_reads(n, u, 18, 32, seed=0x5EED0001)
    (ow, ol), expect = _oracle_counts_of_batch(oracle, b, 0)
    ctr = sq.DeviceCounter(0, expected_unique=2_000_000)      # 2^22 slots = 64 MB > the L2-resident limit
    assert ctr.capacity() == 1 << 22
    if fused:
        arr = ctr.pack_count(b)
        w, l, _ = arr.to_host()
        assert np.array_equal(w, ow) and np.array_equal(l, ol)
    else:
        ctr.insert(sq.pack_batch(b, klass=0))
    keys, counts, _, _ = ctr.export(1)
    kw, kl, _ = keys.to_host()
    assert len(ctr) == len(expect)
    assert counter_dict(kw, kl, counts.cpu().numpy()) == expect
    # a second pass over the same reads doubles every count (table already populated)
    ctr.pack_count(b)
    keys, counts, _, _ = ctr.export(1)
    kw, kl, _ = keys.to_host()
    assert counter_dict(kw, kl, counts.cpu().numpy()) == {k: 2 * v for k, v in expect.items()}


def test_counter_deferred_l2_fallback(sq, oracle, monkeypatch):
    """Tables whose regions do not fit in shared memory (>= 2^30 slots) count the level-1 partitions in order with
    global atomics; SSQ_FORCE_L2_COUNT=1 takes that path on a small table."""
    monkeypatch.setenv("SSQ_FORCE_L2_COUNT", "1")
    n, u = 2_500_000, 1_200_000
    b = sq.synth_reads(n, u, 20, 32, seed=0x5EED0011)
    (ow, ol), expect = _oracle_counts_of_batch(oracle, b, 0)
    ctr = sq.DeviceCounter(0, expected_unique=2_000_000)
    ctr.pack_count(b)
    keys, counts, _, _ = ctr.export(1)
    kw, kl, _ = keys.to_host()
    assert len(ctr) == len(expect)
    assert counter_dict(kw, kl, counts.cpu().numpy()) == expect


@pytest.mark.parametrize("log2_slots", [23, 25])
def test_counter_region_path_table_sizes(sq, oracle, log2_slots):
    """The shared-memory region path with fewer than 256 regions per level-1 partition (tables below 2^28 slots):
    a region then gathers several level-2 partitions."""
    u = (1 << log2_slots) // 2 - 1000
    n = max(3_000_000, (1 << log2_slots) // 3)
    b = sq.synth_reads(n, n // 3, 25, 32, seed=0x5EED0021)
    (ow, ol), expect = _oracle_counts_of_batch(oracle, b, 0)
    ctr = sq.DeviceCounter(0, expected_unique=u)
    assert ctr.capacity() == 1 << log2_slots
    ctr.pack_count(b)
    keys, counts, _, _ = ctr.export(1)
    kw, kl, _ = keys.to_host()
    assert len(ctr) == len(expect)
    assert counter_dict(kw, kl, counts.cpu().numpy()) == expect


def test_counter_bound_exceeded_across_batches(sq, oracle):
    """expected_unique = 40000 gives 2^17 slots; two batches of 70000 distinct keys each do not fit.  The
    table must grow (before the second batch: it already holds more keys than the bound) instead of dropping keys."""
    rng = np.random.default_rng(5)
    ctr = sq.DeviceCounter(0, expected_unique=40_000)
    assert ctr.capacity() == 1 << 17
    total = {}
    for batch in range(3):
        reads = rand_reads(rng, 70_000, 28, 32)
        buf, off = concat(reads)
        ow, ol, _ = oracle.pack_batch(0, buf, off)
        uw, ul, uc, _ = oracle.count(ow, ol, 1)
        for k, v in counter_dict(uw, ul, uc).items():
            total[k] = total.get(k, 0) + v
        ctr.pack_count(buf, off)
        assert len(ctr) == len(total)
    assert ctr.capacity() >= 1 << 19
    keys, counts, _, _ = ctr.export(1)
    kw, kl, _ = keys.to_host()
    assert counter_dict(kw, kl, counts.cpu().numpy()) == total
    # the same through insert() of packed batches, and a bound that grows the table before the pass
    ctr2 = sq.DeviceCounter(0, expected_unique=100_000)            # 2^18 slots; the bound holds, 3 x 70000 reads of 90000 keys
    pool = rand_reads(rng, 90_000, 20, 32)
    tot2 = {}
    for batch in range(3):
        reads = [pool[i] for i in rng.integers(0, len(pool), size=70_000)]
        buf, off = concat(reads)
        ow, ol, _ = oracle.pack_batch(0, buf, off)
        uw, ul, uc, _ = oracle.count(ow, ol, 1)
        for k, v in counter_dict(uw, ul, uc).items():
            tot2[k] = tot2.get(k, 0) + v
        ctr2.insert(sq.pack_batch(buf, off, klass=0))
    assert ctr2.capacity() == 1 << 18
    keys, counts, _, _ = ctr2.export(1)
    kw, kl, _ = keys.to_host()
    assert counter_dict(kw, kl, counts.cpu().numpy()) == tot2


@pytest.mark.parametrize("log2_slots", [28, 29])
def test_counter_bench_geometry(sq, oracle, log2_slots):
    """The geometry bench.py's headline runs on: DeviceCounter(expected_unique = 1e8) -> 2^28 slots =
    65536 regions of 4096 slots, one level-2 partition per region, 6 slices; n >= cap/4 reads so that the deferred path is
    taken; two passes (empty, then populated table) against the C oracle's counts of the same reads.  2^29 slots: the
    8192-slot regions of the next table size."""
    import torch
    from tests.util import canon_counts
    free, _ = torch.cuda.mem_get_info()
    if free < (30 << 30) * (log2_slots - 27):
        pytest.skip("not enough device memory")
    cap = 1 << log2_slots
    n = cap // 4 + 1_000_003                    # >= cap / 4: use_deferred() holds
    u = n // 3                                  # ~ 2.3e7 distinct 32-nt sequences (every key id is its own sequence)
    b = sq.synth_reads(n, u, 32, 32, seed=0x5EED0001)
    buf, off = b.ascii.cpu().numpy(), b.offsets.cpu().numpy()
    ow, ol, _ = oracle.pack_batch(0, buf, off)
    uw, ul, uc, _ = oracle.count(ow, ol, 1)
    ew, el, ec = canon_counts(uw, ul, uc)
    del buf, off, uw, ul, uc
    ctr = sq.DeviceCounter(0, expected_unique=cap // 2 - cap // 8)     # 1.0e8 for 2^28 slots
    assert ctr.capacity() == cap
    arr = ctr.pack_count(b)
    w, l, _ = arr.to_host()
    assert np.array_equal(w, ow) and np.array_equal(l, ol)
    del w, l
    for npass in (1, 2):
        if npass == 2:
            ctr.pack_count(b)                   # second pass: every region is loaded, counted into and written back
        assert len(ctr) == len(ec)
        keys, counts, _, _ = ctr.export(1)
        kw, kl, _ = keys.to_host()
        gw, gl, gc = canon_counts(kw, kl, counts.cpu().numpy())
        assert np.array_equal(gw, ew) and np.array_equal(gl, el), "exported keys differ from the oracle"
        assert np.array_equal(gc, npass * ec), "multiplicities differ from the oracle"
        del keys, counts, kw, kl, gw, gl, gc


@pytest.mark.parametrize("skew", [False, True])
def test_counter_deferred_192(sq, oracle, skew):
    """ShortSeq192 tables too large for L2: records {w0, w1, w2, meta} are scattered to 256 hash partitions by the pack
    kernel and inserted partition by partition.  skew: a third of the reads identical (staging ring and segment overrun)."""
    n, u = 1_500_000, 700_000
    b = sq.synth_reads(n, u, 33, 96, seed=0x5EED0031)
    if skew:
        off = b.offsets.cpu().numpy()
        L0 = int(off[1] - off[0])
        a = b.ascii
        first = a[:L0].clone()
        lens = np.diff(off)
        idx = np.nonzero(lens == L0)[0][::2]                 # same-length reads become copies of read 0
        for i in idx[: n // 3]:
            a[off[i]: off[i] + L0] = first
    (ow, ol), expect = _oracle_counts_of_batch(oracle, b, 1)
    ctr = sq.DeviceCounter(1, expected_unique=1_000_000)       # 2^21 slots x 32 B = 64 MB > the L2-resident limit
    assert ctr.capacity() == 1 << 21
    arr = ctr.pack_count(b)
    w, l, _ = arr.to_host()
    assert np.array_equal(w, ow) and np.array_equal(l, ol)
    keys, counts, _, _ = ctr.export(1)
    kw, kl, _ = keys.to_host()
    assert len(ctr) == len(expect)
    assert counter_dict(kw, kl, counts.cpu().numpy()) == expect
    ctr.pack_count(b)                                          # second pass over a populated table
    keys, counts, _, _ = ctr.export(1)
    kw, kl, _ = keys.to_host()
    assert counter_dict(kw, kl, counts.cpu().numpy()) == {k: 2 * v for k, v in expect.items()}


@pytest.mark.parametrize("L,shift,variant", [(75, 0, "clean"), (33, 5, "clean"), (96, 11, "clean"), (64, 0, "clean"),
                                             (75, 3, "ragged"), (75, 0, "bad")])
def test_counter_deferred_192_uniform_length(sq, oracle, L, shift, variant):
    """ShortSeq192 batches of one read length through the deferred path (the C3 shape), at every alignment class of the
    tile geometry.  shift: the ASCII buffer starts `shift` bytes off a 16-byte boundary.  ragged: two pairs of reads trade
    one base (74 / 76 nt) so that the batch still holds L n bytes but is not uniform.  bad: invalid bases in three reads;
    the lowest one is reported and nothing else changes.  (A fixed-length fast-path kernel that only checked the offsets
    against the progression was built against this test and measured 2 % faster than the general kernel: dropped.)"""
    import torch
    from shortseq_b200 import _lib
    from shortseq_b200.batch import ReadBatch
    n, u = 1_300_000 + 77, 600_000
    b0 = sq.synth_reads(n, u, L, L, seed=0x5EED0077 + L)
    buf, off = b0.ascii.cpu().numpy().copy(), b0.offsets.cpu().numpy().copy()
    if variant == "ragged":
        for i in (1000, 700_001):
            off[i + 1] -= 1            # read i loses its last base to read i + 1
    bad_reads = []
    if variant == "bad":
        bad_reads = [900_000, 412_345, 1_299_999]
        for i in bad_reads:
            buf[off[i] + (i % L)] = ord("N")
    dev = torch.empty(buf.size + 16, dtype=torch.uint8, device=b0.ascii.device)
    dev[shift: shift + buf.size] = torch.from_numpy(buf).to(dev.device)
    b = ReadBatch.make(dev[shift: shift + buf.size], torch.from_numpy(off).to(dev.device))
    ctr = sq.DeviceCounter(1, expected_unique=1_000_000)       # 2^21 slots x 32 B = 64 MB: deferred counting
    arr = ctr.pack_count(b, check=False)
    rep = ctr.ctx.sync()
    if variant == "bad":
        assert rep.code == _lib.ERR_BAD_BASE and rep.first_bad_read == min(bad_reads)
        good = np.ones(n, dtype=bool)
        good[bad_reads] = False
    else:
        assert rep.code == _lib.OK
        good = np.ones(n, dtype=bool)
    clean = buf.copy()
    for i in bad_reads:
        clean[off[i] + (i % L)] = ord("A")                     # the oracle packs a clean copy; bad reads are left out of the counts
    ow, ol, _ = oracle.pack_batch(1, clean, off)
    w, l, _ = arr.to_host()
    assert np.array_equal(l, ol)
    assert np.array_equal(w[good], ow[good])
    uw, ul, uc, _ = oracle.count(ow[good], ol[good], 3)
    expect = counter_dict(uw, ul, uc)
    keys, counts, _, _ = ctr.export(1)
    kw, kl, _ = keys.to_host()
    assert len(ctr) == len(expect)
    assert counter_dict(kw, kl, counts.cpu().numpy()) == expect


@pytest.mark.parametrize("owner_unique", [400_000, 250_000])
def test_counter_merge_regions(sq, oracle, owner_unique):
    """The owner side of the multi-GPU merge on one GPU: P sender tables export one owner's share (tuples + region
    offsets) into receive buffers; ssq_counter_merge_regions counts them region by region in shared memory.
    owner_unique 400k: owner regions = sender regions per owner; 250k: two sender regions per owner region."""
    import torch
    from shortseq_b200 import hashing
    P, me = 4, 1
    dev = None
    senders, expect = [], {}
    for s_ in range(P):
        b = sq.synth_reads(900_000, 1_200_000, 14, 32, seed=0x5EED0081, first_read=900_000 * s_)
        (ow, ol), d = _oracle_counts_of_batch(oracle, b, 0)
        for k, v in d.items():
            expect[k] = expect.get(k, 0) + v
        c = sq.DeviceCounter(0, expected_unique=1_500_000)
        c.pack_count(b)
        senders.append(c)
        dev = c.ctx.device
    assert senders[0].capacity() == 1 << 22 and senders[0].regions() == 1024
    sizes = [c.export_counts(P).cpu().numpy() for c in senders]
    per_owner = senders[0].regions() // P
    total = int(sum(sz[me] for sz in sizes))
    words = torch.empty(total, dtype=torch.int64, device=dev)
    lens = torch.empty(total, dtype=torch.uint8, device=dev)
    counts = torch.empty(total, dtype=torch.int64, device=dev)
    rb = torch.empty((P, per_owner + 1), dtype=torch.int64, device=dev)
    big = int(max(sz.max() for sz in sizes))
    dump_w = torch.empty(big, dtype=torch.int64, device=dev); dump_l = torch.empty(big, dtype=torch.uint8, device=dev)
    dump_c = torch.empty(big, dtype=torch.int64, device=dev); dump_rb = torch.empty(per_owner + 1, dtype=torch.int64, device=dev)
    at = 0
    for s_, c in enumerate(senders):
        table = np.empty((3, P), dtype=np.int64)
        for p_ in range(P):
            mine = p_ == me
            table[0, p_] = words.data_ptr() + 8 * at if mine else dump_w.data_ptr()
            table[1, p_] = lens.data_ptr() + at if mine else dump_l.data_ptr()
            table[2, p_] = counts.data_ptr() + 8 * at if mine else dump_c.data_ptr()
        c.export_to(P, torch.from_numpy(table).to(dev), first_part=(s_ + 1) % P)
        rbt = np.array([rb[s_].data_ptr() if p_ == me else dump_rb.data_ptr() for p_ in range(P)], dtype=np.int64)
        c.export_region_bases(P, torch.from_numpy(rbt).to(dev))
        at += int(sizes[s_][me])
    torch.cuda.synchronize()
    owner = sq.DeviceCounter(0, expected_unique=owner_unique, hash_rot=2)
    owner.merge_regions_raw(words.data_ptr(), lens.data_ptr(), counts.data_ptr(), total, [int(sz[me]) for sz in sizes],
                            [per_owner] * P, rb.data_ptr(), per_owner + 1)
    keys, cnt, _, _ = owner.export(1)
    kw, kl, _ = keys.to_host()
    got = counter_dict(kw, kl, cnt.cpu().numpy())
    ew = np.array([k[1][0] for k in expect], dtype=np.uint64)
    el = np.array([k[0] for k in expect], dtype=np.uint8)
    own = hashing.owner_rank(ew, el, 0, P)
    mine_expect = {k: v for (k, v), o in zip(expect.items(), own) if o == me}
    assert len(owner) == len(mine_expect)
    assert got == mine_expect
    # a second merge into the now populated regions (loaded, only the touched slots written back), with weights whose
    # low halves overflow 32 bits when added up: the carry into the high half of the 64-bit shared-memory deltas
    big_w = (1 << 31) + 12345
    counts2 = counts * big_w
    owner.merge_regions_raw(words.data_ptr(), lens.data_ptr(), counts2.data_ptr(), total, [int(sz[me]) for sz in sizes],
                            [per_owner] * P, rb.data_ptr(), per_owner + 1)
    keys, cnt, _, _ = owner.export(1)
    kw, kl, _ = keys.to_host()
    assert len(owner) == len(mine_expect)
    assert counter_dict(kw, kl, cnt.cpu().numpy()) == {k: v * (1 + big_w) for k, v in mine_expect.items()}


@pytest.mark.parametrize("klass", [0, 1])
def test_counter_merge_exports(sq, oracle, klass):
    """ssq_counter_merge of the concatenated exports of several counters equals the counts of all their reads."""
    import torch
    lo, hi = (10, 32) if klass == 0 else (33, 96)
    batches = [sq.synth_reads(120_000 + 7_000 * j, 30_000, lo, hi, seed=0x5EED0071, first_read=1_000_000 * j) for j in range(3)]
    expect = {}
    ws, ls, cs, sizes = [], [], [], []
    for b in batches:
        _, d = _oracle_counts_of_batch(oracle, b, klass)
        for k, v in d.items():
            expect[k] = expect.get(k, 0) + v
        ctr = sq.DeviceCounter(klass, expected_unique=40_000)
        ctr.pack_count(b)
        keys, counts, _, _ = ctr.export(1)
        ws.append(keys.words); ls.append(keys.lens); cs.append(counts); sizes.append(len(keys))
    w, l, c = torch.cat(ws).contiguous(), torch.cat(ls).contiguous(), torch.cat(cs).contiguous()
    owner = sq.DeviceCounter(klass, expected_unique=60_000)
    owner.merge_raw(w.data_ptr(), l.data_ptr(), c.data_ptr(), int(l.numel()))
    keys, counts, _, _ = owner.export(1)
    kw, kl, _ = keys.to_host()
    assert counter_dict(kw, kl, counts.cpu().numpy()) == expect


@pytest.mark.parametrize("klass", [0, 1])
def test_counter_deferred_tight_segments(sq, oracle, klass, monkeypatch):
    """Partition segments sized with no slack at all: about half of them overflow by a few keys, so every way a segment
    can end (full lines up to the cap, a partial last line that fits or does not) occurs; the excess is inserted
    directly and the counts stay exact."""
    monkeypatch.setenv("SSQ_SEG_SLACK_PCT", "0")
    monkeypatch.setenv("SSQ_SEG_SLACK_ABS", "0")
    n, u = (3_000_000, 1_400_000) if klass == 0 else (1_500_000, 700_000)
    lo, hi = (20, 32) if klass == 0 else (33, 96)
    b = sq.synth_reads(n, u, lo, hi, seed=0x5EED0051)
    (ow, ol), expect = _oracle_counts_of_batch(oracle, b, klass)
    ctr = sq.DeviceCounter(klass, expected_unique=2_000_000 if klass == 0 else 1_000_000)
    # a first pass over other reads leaves stale keys in the partition buffers: entries past a segment's real end
    # must never be read back
    ctr.pack_count(sq.synth_reads(n, u, lo, hi, seed=0x5EED0052))
    ctr.clear()
    ctr.pack_count(b)
    keys, counts, _, _ = ctr.export(1)
    kw, kl, _ = keys.to_host()
    assert len(ctr) == len(expect)
    assert counter_dict(kw, kl, counts.cpu().numpy()) == expect


def test_counter_deferred_partition_overflow_and_growth(sq, oracle):
    """Half of the reads are one sequence: its hash partition overflows its buffer and the excess is inserted
    directly; the distinct keys exceed 60 % of the table, so it grows after the pass."""
    n, u = 6_000_000, 50_000_000
    b = sq.synth_reads(n, u, 32, 32, seed=0x5EED0009)
    a = b.ascii.view(n, 32)
    a[::2] = a[0].clone()                       # every second read becomes a copy of read 0
    (ow, ol), expect = _oracle_counts_of_batch(oracle, b, 0)
    ctr = sq.DeviceCounter(0, expected_unique=2_000_000)
    ctr.pack_count(b)
    assert ctr.capacity() == 1 << 23            # grew
    keys, counts, _, _ = ctr.export(1)
    kw, kl, _ = keys.to_host()
    got = counter_dict(kw, kl, counts.cpu().numpy())
    assert got == expect and max(got.values()) >= n // 2


def test_host_pipeline_lens_variant(sq, oracle):
    """ssq_host_pack_count_lens: boundaries as one uint8 length per read; same words and counts."""
    import ctypes as C
    from shortseq_b200 import _lib
    rng = np.random.default_rng(112)
    for klass, lo, hi in ((0, 0, 32), (1, 33, 96)):
        pool = rand_reads(rng, 2000, lo, hi)
        reads = [pool[i] for i in rng.integers(0, 2000, size=60_000)]
        buf, off = concat(reads)
        lens_in = np.diff(off).astype(np.uint8)
        ctr = sq.DeviceCounter(klass, expected_unique=3000)
        words = np.zeros(len(reads) if klass == 0 else (len(reads), 3), np.uint64)
        rep = _lib.Report()
        _lib.check(_lib.lib().ssq_host_pack_count_lens(ctr.ctx.bind(), ctr.handle, buf.ctypes.data, lens_in.ctypes.data, len(reads),
                                                      words.ctypes.data, 7001, C.byref(rep)))
        assert rep.code == 0
        ow, ol, _ = oracle.pack_batch(klass, buf, off)
        assert np.array_equal(words, ow)
        uw, ul, uc, _ = oracle.count(ow, ol, 3 if klass == 1 else 1)
        keys, counts, _, _ = ctr.export(1)
        kw, kl, _ = keys.to_host()
        assert counter_dict(kw, kl, counts.cpu().numpy()) == counter_dict(uw, ul, uc)
    # a bad base is reported with its index in the whole batch, not in its chunk
    reads = [b"ACGT" * 5] * 20_000
    reads[15_000] = b"ACGTNACGTACGTACGTACG"
    buf, off = concat(reads)
    lens_in = np.diff(off).astype(np.uint8)
    ctr = sq.DeviceCounter(0, expected_unique=10)
    rep = _lib.Report()
    _lib.check(_lib.lib().ssq_host_pack_count_lens(ctr.ctx.bind(), ctr.handle, buf.ctypes.data, lens_in.ctypes.data, len(reads), None, 4096,
                                                  C.byref(rep)))
    assert rep.code == _lib.ERR_BAD_BASE and rep.first_bad_read == 15_000
