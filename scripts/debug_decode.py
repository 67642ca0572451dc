import sys, torch
import shortseq_b200 as sq
for n in [1_000_000, 10_000_000, 40_000_000, 70_000_000, 134_000_000]:
    b = sq.synth_reads(n, max(1, n // 10), 32, 32)
    arr = sq.pack_batch(b, klass=0)
    print("packed", n, flush=True)
    try:
        a, o = arr.decode()
        torch.cuda.synchronize()
        print("decode ok", n, int(o[-1]), bool((a == b.ascii).all()), flush=True)
    except Exception as e:
        print("decode FAIL", n, str(e)[:200], flush=True)
        break
    d = sq.hamming_batch(arr, arr); torch.cuda.synchronize()
    print("hamming ok", int(d.max()), flush=True)
