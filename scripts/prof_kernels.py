"""Runs each kernel a few times at a mid size (for ncu captures: -k regex:<name> -s 1 -c 1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import shortseq_b200 as sq
from shortseq_b200 import _lib
from shortseq_b200._runtime import ptr

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 26
L = int(sys.argv[2]) if len(sys.argv) > 2 else 32
klass = 0 if L <= 32 else 1
b = sq.synth_reads(n, n // 10, L, L)
b2 = sq.synth_reads(n, n // 10, L, L, seed=77)
ctx = b.ctx; lib = _lib.lib(); h = ctx.bind()
for _ in range(3):
    arr = sq.pack_batch(b, klass=klass)
arr2 = sq.pack_batch(b2, klass=klass)
for _ in range(3):
    out = arr.decode()
for _ in range(3):
    d = sq.hamming_batch(arr, arr2)
refs = sq.pack_batch(sq.synth_reads(256, 256, L, L, seed=5), klass=klass)
for _ in range(2):
    sq.hamming_refset(arr, refs, thresh=1)
ctr = sq.DeviceCounter(klass, expected_unique=n // 10)
ctr.insert(arr)
for _ in range(2):
    ctr.export(8)
torch.cuda.synchronize()
print("ok", n, L)
