"""Development: weighted ShortSeq192 insert (owner side of a C3 merge) -- P copies of a table's export merged into an owner table."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import shortseq_b200 as sq
from shortseq_b200 import _lib
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2 * 10**8
u = int(float(sys.argv[2])) if len(sys.argv) > 2 else 31_250_000
P = int(sys.argv[3]) if len(sys.argv) > 3 else 4
b = sq.synth_reads(n, u, 75, 75)
local = sq.DeviceCounter(1, expected_unique=u)
local.pack_count(b)
del b
keys, counts, _, parts = local.export(P)
n0 = int(parts[0])
w = keys.words[:n0].repeat(P, 1).contiguous(); l = keys.lens[:n0].repeat(P).contiguous(); c = counts[:n0].repeat(P).contiguous()
owner = sq.DeviceCounter(1, expected_unique=int(1.1 * u / P) + 1024, hash_rot=P.bit_length() - 1)
def merge():
    _lib.lib().ssq_counter_clear(owner.handle); owner.merge(w, l, c)
merge(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3): merge()
torch.cuda.synchronize()
print(f"P={P}: weighted insert of {len(l)} ShortSeq192 tuples into cap={owner.capacity()}: {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms (incl. clear) -> {len(owner)} keys (expected {n0})")
