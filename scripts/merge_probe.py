import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
"""Times the pieces of the multi-GPU merge on one GPU: export grouped by owner, weighted merge of one owner's share."""
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
owner = sq.DeviceCounter(0, expected_unique=2 * u // P, hash_rot=rot)
from shortseq_b200 import _lib
def merge():
    _lib.lib().ssq_counter_clear(owner.handle); owner.merge(w, l, c)
ms, _ = t(merge)
print(f"merge of {len(l)} tuples into owner table cap={owner.capacity()}: {ms:.2f} ms -> {len(owner)} keys")
def merge_b():
    _lib.lib().ssq_counter_clear(owner.handle); owner.merge_raw(w.data_ptr(), l.data_ptr(), c.data_ptr(), len(l), block_counts=[len(l) // P] * P)
ms, _ = t(merge_b)
print(f"merge_blocks (lockstep over the {P} blocks): {ms:.2f} ms -> {len(owner)} keys")
# shuffled order for comparison
perm = torch.randperm(len(l), device="cuda")
w2, l2, c2 = w[perm].contiguous(), l[perm].contiguous(), c[perm].contiguous()
def merge2():
    _lib.lib().ssq_counter_clear(owner.handle); owner.merge(w2, l2, c2)
ms, _ = t(merge2)
print(f"merge (shuffled order): {ms:.2f} ms")
