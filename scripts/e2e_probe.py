import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C, time, sys, torch
import shortseq_b200 as sq
from shortseq_b200 import _lib
n = 1 << 27; L = 32
b = sq.synth_reads(n, n // 10, L, L)
ctx = b.ctx; lib = _lib.lib(); h = ctx.bind()
h_ascii = torch.empty(n * L, dtype=torch.uint8).pin_memory(); h_off = torch.empty(n + 1, dtype=torch.int64).pin_memory()
h_ascii.copy_(b.ascii[: n * L]); h_off.copy_(b.offsets)
h_words = torch.empty(n, dtype=torch.int64).pin_memory(); h_lens = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n * L, dtype=torch.uint8, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(h_ascii, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"raw H2D {n*L/dt/1e9:.1f} GB/s")
dw = torch.empty(n, dtype=torch.int64, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); h_words.copy_(dw, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"raw D2H {n*8/dt/1e9:.1f} GB/s")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(h_ascii, non_blocking=True)
    s2 = torch.cuda.Stream()
    with torch.cuda.stream(s2): h_words.copy_(dw, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"concurrent H2D {n*L/dt/1e9:.1f} GB/s + D2H {n*8/dt/1e9:.1f} GB/s")
del d, dw
ctr = sq.DeviceCounter(0, expected_unique=n // 10)
rep = _lib.Report()
for chunk in (1 << 20, 1 << 22, 1 << 24, 1 << 25):
    for outs in (True, False):
        ts = []
        for _ in range(3):
            lib.ssq_counter_clear(ctr.handle); torch.cuda.synchronize(); t0 = time.perf_counter()
            _lib.check(lib.ssq_host_pack_count(h, ctr.handle, h_ascii.data_ptr(), h_off.data_ptr(), n, h_words.data_ptr() if outs else None,
                                               h_lens.data_ptr() if outs else None, chunk, C.byref(rep)))
            ts.append(time.perf_counter() - t0)
        dt = min(ts)
        print(f"chunk {chunk:>9} outputs={outs}: {dt*1e3:7.1f} ms  {n*L/dt/1e9:6.1f} Gbases/s  H2D {(n*L+8*n)/dt/1e9:5.1f} GB/s")

h_l8 = torch.full((n,), L, dtype=torch.uint8).pin_memory()
for chunk in (1 << 22, 1 << 23):
    for outs in (True, False):
        ts = []
        for _ in range(3):
            lib.ssq_counter_clear(ctr.handle); torch.cuda.synchronize(); t0 = time.perf_counter()
            _lib.check(lib.ssq_host_pack_count_lens(h, ctr.handle, h_ascii.data_ptr(), h_l8.data_ptr(), n, h_words.data_ptr() if outs else None, chunk, C.byref(rep)))
            ts.append(time.perf_counter() - t0)
        dt = min(ts)
        print(f"LENS chunk {chunk:>9} outputs={outs}: {dt*1e3:7.1f} ms  {n*L/dt/1e9:6.1f} Gbases/s  H2D {(n*L+n)/dt/1e9:5.1f} GB/s")
