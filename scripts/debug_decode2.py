import sys, torch, ctypes as C
import shortseq_b200 as sq
from shortseq_b200 import _lib
from shortseq_b200._runtime import ptr
n = int(sys.argv[1])
b = sq.synth_reads(n, max(1, n // 10), 32, 32)
arr = sq.pack_batch(b, klass=0)
print("packed", n, "lens min/max", int(arr.lens.min()), int(arr.lens.max()), flush=True)
ctx = arr.ctx; L = _lib.lib(); h = ctx.bind()
out_off = ctx.empty((n + 1,), torch.int64)
_lib.check(L.ssq_lens_to_offsets(h, ptr(arr.lens), 1, n, ptr(out_off)))
torch.cuda.synchronize()
ref = torch.zeros(n + 1, dtype=torch.int64, device="cuda"); ref[1:] = torch.cumsum(arr.lens.to(torch.int64), 0)
bad = (ref != out_off).nonzero()
print("scan mismatches", bad.numel(), bad[:5].flatten().tolist(), out_off[-3:].tolist(), flush=True)
out = torch.zeros(32 * n + (1 << 20), dtype=torch.uint8, device="cuda")
_lib.check(L.ssq_decode64(h, ptr(arr.words), ptr(arr.lens), n, ptr(ref), ptr(out)))
torch.cuda.synchronize()
print("decode ok", bool((out[:32*n] == b.ascii).all()), int(out[32*n:].max()), flush=True)
