"""Quick device timing of the main kernels (development aid, not the bench contract)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import sys
import time

import torch

import shortseq_b200 as sq
from shortseq_b200 import _lib
from shortseq_b200._runtime import ptr


def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(iters):
        fn()
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / iters


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 27
    u = int(float(sys.argv[2])) if len(sys.argv) > 2 else n // 10
    L = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    klass = 0 if L <= 32 else 1
    W = 1 if klass == 0 else 3
    t0 = time.time()
    b = sq.synth_reads(n, u, L, L)
    torch.cuda.synchronize()
    print(f"synth {n} x {L}nt, U={u}: {time.time()-t0:.2f}s")
    ctx = b.ctx
    lib = _lib.lib()
    words = ctx.empty((n,) if W == 1 else (n, 3), torch.int64)
    lens = ctx.empty((n,), torch.uint8)
    h = ctx.bind()
    nbytes = int(b.ascii.numel())
    packfn = lib.ssq_pack64 if klass == 0 else lib.ssq_pack192
    ms = timeit(lambda: _lib.check(packfn(h, ptr(b.ascii), nbytes, ptr(b.offsets), n, ptr(words), ptr(lens))))
    bytes_pack = n * (L + 8 + 8 * W + 1)
    print(f"pack:        {ms:8.3f} ms  {n*L/ms/1e6:8.1f} Gbases/s  {bytes_pack/ms/1e6:8.1f} GB/s algorithmic")
    ctr = sq.DeviceCounter(klass, expected_unique=u)

    def ins():
        lib.ssq_counter_clear(ctr.handle)
        _lib.check(lib.ssq_counter_insert(ctr.handle, ptr(words), ptr(lens), n))
    ms = timeit(ins, 3)
    print(f"insert:      {ms:8.3f} ms  ({n/ms/1e6:.2f} Greads/s) uniques={len(ctr)} cap={ctr.capacity()}")

    def pc():
        lib.ssq_counter_clear(ctr.handle)
        _lib.check(lib.ssq_counter_pack_count(ctr.handle, ptr(b.ascii), nbytes, ptr(b.offsets), n, ptr(words), ptr(lens)))
    ms = timeit(pc, 3)
    nu = len(ctr)
    bytes_pc = bytes_pack + nu * (8 * W + 9)
    p1, p2 = C.c_float(), C.c_float()
    lib.ssq_counter_last_pass_ms(ctr.handle, C.byref(p1), C.byref(p2))
    d = [C.c_float(), C.c_float(), C.c_float()]
    lib.ssq_counter_last_pass_detail(ctr.handle, C.byref(d[0]), C.byref(d[1]), C.byref(d[2]))
    print(f"pack+count:  {ms:8.3f} ms  {n*L/ms/1e6:8.1f} Gbases/s  {bytes_pc/ms/1e6:8.1f} GB/s algorithmic  uniques={nu}  phase1={p1.value:.3f} ms phase2={p2.value:.3f} ms"
          f"  (pack+scatter {d[0].value:.3f}, region scatter {d[1].value:.3f}, region count {d[2].value:.3f})")
    rep = ctx.sync()
    print("report", rep.code, rep.first_bad_read)
    arr = sq.ShortSeqArray(ctx, klass, words, lens)
    ms = timeit(lambda: arr.decode(), 3)
    print(f"decode(+scan): {ms:8.3f} ms  {n*L/ms/1e6:8.1f} Gbases/s")
    ms = timeit(lambda: sq.hamming_batch(arr, arr), 3)
    print(f"hamming:     {ms:8.3f} ms  {n/ms/1e6:8.2f} Gpairs/s")
    t0 = time.time()
    keys, counts, _, parts = ctr.export(8)
    torch.cuda.synchronize()
    print(f"export(8 parts): {time.time()-t0:.3f}s parts={parts.tolist()}")


if __name__ == "__main__":
    main()
