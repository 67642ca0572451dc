"""Small runs of the round-2 kernels for compute-sanitizer (memcheck / racecheck): ShortSeqVar pack + decode (bulk-copy
pipeline, producer / consumer warps), ShortSeq192 deferred count, the single-object calls, ShortSeq64 deferred count."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import shortseq_b200 as sq
from oracle import oracle as O

buf, off = O.synth_reads(0x5EED0002, 0, 3000, 3000, 97, 1024)
arr = sq.pack_batch(buf, off, klass=sq.CLASS_VAR)
w, l, wo = arr.to_host()
ow, ol, oo = O.pack_batch(sq.CLASS_VAR, buf, off)
assert np.array_equal(w, ow) and np.array_equal(wo, oo)
out, out_off = arr.decode()
assert np.array_equal(out.cpu().numpy(), buf)
for klass, lo, hi, n, u in ((sq.CLASS_192, 33, 96, 3_000_000, 1_200_000), (sq.CLASS_64, 32, 32, 6_000_000, 3_000_000), (sq.CLASS_64, 10, 32, 6_000_000, 3_000_000)):
    b = sq.synth_reads(n, u, lo, hi)
    ctr = sq.DeviceCounter(klass, expected_unique=u)
    ctr.pack_count(b)
    keys, counts, _, _ = ctr.export(2)
    assert int(counts.sum().item()) == n, (klass, int(counts.sum().item()))
a, b2 = sq.pack("ACGT" * 20), sq.pack("ACGA" * 20)
assert str(a) == "ACGT" * 20 and (a ^ b2) == 20
torch.cuda.synchronize()
print("sanitize_small ok")
