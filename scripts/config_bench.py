"""Device timings of the BASELINE.json configs that are not the bench line (development aid): ShortSeqVar pack + decode
round trip, batched Hamming (pairs and against a reference set), ShortSeq64/192 decode.  Algorithmic bytes per SURVEY 8d."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import shortseq_b200 as sq


def timeit(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def line(name, ms, bases, nbytes):
    print(f"{name:44s} {ms:9.3f} ms  {bases/ms/1e6:9.1f} Gbases/s  {nbytes/ms/1e6:8.1f} GB/s algorithmic ({nbytes/ms/1e6/6549.1*100:4.1f}% of HBM peak)", flush=True)


def main():
    # C4: variable-length reads 97..1024 nt
    n = 20_000_000
    b = sq.synth_reads(n, n // 10, 97, 1024, seed=0x5EED0041)
    total = int(b.ascii.numel())
    arr = sq.pack_batch(b, klass=sq.CLASS_VAR)
    nwords = int(arr.words.numel())
    ms = timeit(lambda: sq.pack_batch(b, klass=sq.CLASS_VAR))
    line(f"ShortSeqVar pack ({n/1e6:.0f}M reads, mean {total/n:.0f} nt)", ms, total, total + 8 * n + 8 * nwords + 8 * n + 2 * n)
    ms = timeit(lambda: arr.decode())
    line("ShortSeqVar decode (+ offsets scan)", ms, total, 8 * nwords + 8 * n + 2 * n + total + 8 * n)
    out, off = arr.decode()
    assert torch.equal(out, b.ascii) and torch.equal(off, b.offsets), "decode(pack(x)) != x"
    del b, arr, out, off
    torch.cuda.empty_cache()
    # decode of fixed classes
    for klass, L, n in ((sq.CLASS_64, 32, 500_000_000), (sq.CLASS_192, 75, 200_000_000)):
        b = sq.synth_reads(n, n // 10, L, L)
        arr = sq.pack_batch(b, klass=klass)
        W = 1 if klass == sq.CLASS_64 else 3
        ms = timeit(lambda: arr.decode())
        line(f"ShortSeq{'64' if W == 1 else '192'} decode ({n/1e6:.0f}M x {L} nt, + offsets scan)", ms, n * L, n * (8 * W + 1 + 8 + L))
        del b, arr
        torch.cuda.empty_cache()
    # C5: Hamming, UMI collapse style
    for klass, L, n in ((sq.CLASS_64, 12, 100_000_000), (sq.CLASS_192, 96, 100_000_000)):
        W = 1 if klass == sq.CLASS_64 else 3
        a = sq.pack_batch(sq.synth_reads(n, n // 10, L, L, seed=1), klass=klass)
        c = sq.pack_batch(sq.synth_reads(n, n // 10, L, L, seed=2), klass=klass)
        ms = timeit(lambda: sq.hamming_batch(a, c))
        line(f"Hamming pairs ({n/1e6:.0f}M x {L} nt)", ms, n * L, n * (2 * (8 * W + 1) + 4))
        for R in (64, 1024):
            refs = sq.pack_batch(sq.synth_reads(R, R, L, L, seed=3), klass=klass)
            ms = timeit(lambda: sq.hamming_refset(a, refs, thresh=1), iters=2)
            print(f"Hamming vs reference set ({n/1e6:.0f}M x {L} nt, R={R:5d})  {ms:9.3f} ms  {n*R/ms/1e9:8.2f} Tcomparisons/s", flush=True)
        del a, c
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
