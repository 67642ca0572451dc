"""Runs the fused pack+count a few times (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import shortseq_b200 as sq
from shortseq_b200 import _lib
from shortseq_b200._runtime import ptr

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 26
u = int(float(sys.argv[2])) if len(sys.argv) > 2 else n // 10
L = int(sys.argv[3]) if len(sys.argv) > 3 else 32
klass = 0 if L <= 32 else 1
b = sq.synth_reads(n, u, L, L)
ctx = b.ctx; lib = _lib.lib(); h = ctx.bind()
words = ctx.empty((n,) if klass == 0 else (n, 3), torch.int64)
lens = ctx.empty((n,), torch.uint8)
ctr = sq.DeviceCounter(klass, expected_unique=u)
for _ in range(3):
    lib.ssq_counter_clear(ctr.handle)
    _lib.check(lib.ssq_counter_pack_count(ctr.handle, ptr(b.ascii), int(b.ascii.numel()), ptr(b.offsets), n, ptr(words), ptr(lens)))
torch.cuda.synchronize()
import ctypes as C
d = [C.c_float(), C.c_float(), C.c_float()]
lib.ssq_counter_last_pass_detail(ctr.handle, C.byref(d[0]), C.byref(d[1]), C.byref(d[2]))
print("ok", n, u, L, len(ctr), "pack+scatter %.3f ms, region scatter %.3f ms, count %.3f ms" % (d[0].value, d[1].value, d[2].value),
      os.environ.get("SSQ_LIB", ""))
