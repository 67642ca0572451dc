"""Consistency of the deferred counting paths on tables beyond 2^28 slots (development aid): 2^29 slots = 8192-slot
regions counted in shared memory, 2^30 slots = the L2-ordered fallback; both against a direct-insert counter."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time
import torch
import shortseq_b200 as sq


def digest(ctr):
    keys, counts, _, _ = ctr.export(1)
    w = keys.words.view(-1)
    # order-independent digest of (word, len, count)
    h = (w * 0x9E3779B97F4A7C15 + keys.lens.to(torch.int64) * 0x632BE59BD9B4E019 + counts * 0x2545F4914F6CDD1D)
    return int(h.sum().item()), int(counts.sum().item()), len(ctr)


for expected, n in ((200_000_000, 150_000_000), (400_000_000, 280_000_000)):
    b = sq.synth_reads(n, n // 8, 28, 32, seed=0x5EED0061)
    big = sq.DeviceCounter(0, expected_unique=expected)
    t0 = time.perf_counter()
    big.pack_count(b)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    d1 = digest(big)
    cap = big.capacity()
    del big
    torch.cuda.empty_cache()
    ref = sq.DeviceCounter(0, expected_unique=n // 8)          # 2^25..2^27 slots: the usual region path
    ref.pack_count(b)
    d2 = digest(ref)
    print(f"table 2^{cap.bit_length()-1} slots, {n} reads: {dt*1e3:.1f} ms  digest {'OK' if d1 == d2 else 'MISMATCH'} {d1} {d2}", flush=True)
    del ref, b
    torch.cuda.empty_cache()
