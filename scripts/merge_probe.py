"""Times the pieces of the multi-GPU merge on one GPU: export grouped by owner, weighted merge of one owner's share."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys, time, torch
import shortseq_b200 as sq
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**9
u = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10**8
P = int(sys.argv[3]) if len(sys.argv) > 3 else 8
b = sq.synth_reads(n, u, 32, 32)
local = sq.DeviceCounter(0, expected_unique=u)
local.pack_count(b)
del b
torch.cuda.synchronize()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3, r
ms, (keys, counts, _, parts) = t(lambda: local.export(P))
print(f"export({P}): {ms:.2f} ms for {len(keys)} uniques, parts {parts.tolist()[:3]}...")
parts = parts.cpu().numpy()
# what one owner receives from P sources: here P copies of this rank's share for owner 0 (same keys, like the bench)
sl = slice(0, int(parts[0]))
w = keys.words[sl].repeat(P); l = keys.lens[sl].repeat(P); c = counts[sl].repeat(P)
rot = P.bit_length() - 1
owner = sq.DeviceCounter(0, expected_unique=int(1.1 * u / P) + 1024, hash_rot=rot)   # like bench.py
from shortseq_b200 import _lib
def merge():
    _lib.lib().ssq_counter_clear(owner.handle); owner.merge(w, l, c)
ms, _ = t(merge)
print(f"merge of {len(l)} tuples into owner table cap={owner.capacity()}: {ms:.2f} ms -> {len(owner)} keys")
# region-aligned merge: P senders = P exports of the local table's partition 0 with their region offsets
import numpy as np
per_owner = local.regions() // P
n0 = int(parts[0])
W = torch.empty(P * n0, dtype=torch.int64, device="cuda"); L = torch.empty(P * n0, dtype=torch.uint8, device="cuda"); Cn = torch.empty(P * n0, dtype=torch.int64, device="cuda")
rb = torch.empty((P, per_owner + 1), dtype=torch.int64, device="cuda")
big = int(parts.max())
dw = torch.empty(big, dtype=torch.int64, device="cuda"); dl = torch.empty(big, dtype=torch.uint8, device="cuda"); dc = torch.empty(big, dtype=torch.int64, device="cuda"); drb = torch.empty(per_owner + 1, dtype=torch.int64, device="cuda")
for s_ in range(P):
    tb = np.empty((3, P), dtype=np.int64)
    for p_ in range(P):
        tb[0, p_] = W.data_ptr() + 8 * s_ * n0 if p_ == 0 else dw.data_ptr()
        tb[1, p_] = L.data_ptr() + s_ * n0 if p_ == 0 else dl.data_ptr()
        tb[2, p_] = Cn.data_ptr() + 8 * s_ * n0 if p_ == 0 else dc.data_ptr()
    local.export_to(P, torch.from_numpy(tb).cuda())
    local.export_region_bases(P, torch.from_numpy(np.array([rb[s_].data_ptr() if p_ == 0 else drb.data_ptr() for p_ in range(P)], dtype=np.int64)).cuda())
torch.cuda.synchronize()
def merge_r():
    _lib.lib().ssq_counter_clear(owner.handle)
    owner.merge_regions_raw(W.data_ptr(), L.data_ptr(), Cn.data_ptr(), P * n0, [n0] * P, [per_owner] * P, rb.data_ptr(), per_owner + 1)
ms, _ = t(merge_r)
print(f"merge_regions (shared-memory, region by region): {ms:.2f} ms -> {len(owner)} keys")
# shuffled order for comparison
perm = torch.randperm(len(l), device="cuda")
w2, l2, c2 = w[perm].contiguous(), l[perm].contiguous(), c[perm].contiguous()
def merge2():
    _lib.lib().ssq_counter_clear(owner.handle); owner.merge(w2, l2, c2)
ms, _ = t(merge2)
print(f"merge (shuffled order): {ms:.2f} ms")
