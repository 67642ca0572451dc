"""End-to-end timing of the GPU FASTQ ingest (host text buffer -> counts) next to the reference's read_and_count_fastq
on a smaller file (development aid).  usage: fastq_bench.py [n_reads] [n_distinct] [read_len]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import time

import numpy as np
import torch

import shortseq_b200 as sq
from shortseq_b200 import _lib


def make_fastq(n, u, L, seed=1):
    rng = np.random.default_rng(seed)
    pool = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(u, L))
    rec = 10 + (L + 1) + 2 + (L + 1)
    a = np.empty((n, rec), dtype=np.uint8)
    a[:, 0] = ord("@")
    ids = np.arange(n, dtype=np.int64)
    for d in range(8):
        a[:, 8 - d] = ord("0") + (ids // 10 ** d) % 10
    a[:, 9] = 10
    a[:, 10:10 + L] = pool[rng.integers(0, u, size=n)]
    a[:, 10 + L] = 10
    a[:, 11 + L] = ord("+")
    a[:, 12 + L] = 10
    a[:, 13 + L:13 + 2 * L] = 73
    a[:, 13 + 2 * L] = 10
    return a.reshape(-1)


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
    u = int(float(sys.argv[2])) if len(sys.argv) > 2 else n // 10
    L = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    text = make_fastq(n, u, L)
    print(f"FASTQ: {n} reads x {L} nt, {u} distinct, {text.size/1e9:.2f} GB", flush=True)
    pinned = torch.empty(text.size, dtype=torch.uint8).pin_memory()
    pinned.numpy()[:] = text
    ctx = sq.DeviceCounter(0, expected_unique=16).ctx
    lib, h = _lib.lib(), ctx.bind()
    klass = 0 if L <= 32 else 1
    for label, ptr_ in (("pinned host buffer", pinned.data_ptr()), ("pageable host buffer", text.ctypes.data)):
        best = None
        for _ in range(3):
            ctr = sq.DeviceCounter(klass, expected_unique=u)
            other = sq.DeviceCounter(1 - klass, expected_unique=16)
            c64, c192 = (ctr, other) if klass == 0 else (other, ctr)
            nr, nl, fl, rep = C.c_int64(), C.c_int64(), C.c_int64(), _lib.Report()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _lib.check(lib.ssq_host_fastq_count(h, c64.handle, c192.handle, ptr_, int(text.size), 0, 0, C.byref(nr), C.byref(nl),
                                                C.byref(fl), C.byref(rep)))
            uniq = len(ctr)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            assert rep.code == 0 and nr.value == n
        print(f"ssq_host_fastq_count, {label}: {best*1e3:8.1f} ms  {text.size/best/1e9:6.1f} GB/s of FASTQ  {n*L/best/1e9:6.2f} Gbases/s  uniques={uniq}", flush=True)
    # the reference on a slice of the same file
    from oracle import ref as R
    ref = R.load()
    if ref is not None:
        m = min(n, 2_000_000)
        rec = text.size // n
        path = "/tmp/ssq_bench.fastq"
        text[: m * rec].tofile(path)
        t0 = time.perf_counter()
        c = ref.read_and_count_fastq(path)
        dt = time.perf_counter() - t0
        print(f"reference read_and_count_fastq ({m} reads, 1 thread): {dt*1e3:8.1f} ms  {m*rec/dt/1e9:6.3f} GB/s of FASTQ  {m*L/dt/1e9:6.3f} Gbases/s  uniques={len(c)}")
        t0 = time.perf_counter()
        g = sq.read_and_count_fastq(path)
        dt = time.perf_counter() - t0
        print(f"shortseq_b200.read_and_count_fastq (same file, incl. boxing {len(g)} keys into Python objects): {dt*1e3:8.1f} ms")
        os.remove(path)


if __name__ == "__main__":
    main()
