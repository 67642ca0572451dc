"""Per-call wall time of the drop-in's sq.pack / str / ^ (batches of one on the GPU) next to the reference's, when the
built reference is present.  usage: percall.py [calls]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shortseq_b200 as sq


def bench(mod, n):
    import random
    rng = random.Random(7)
    seqs = ["".join(rng.choice("ACGT") for _ in range(L)) for L in (22, 75, 300) for _ in range(n // 3)]
    t = time.perf_counter(); objs = [mod.pack(s) for s in seqs]; tp = (time.perf_counter() - t) / len(seqs)
    t = time.perf_counter(); strs = [str(o) for o in objs]; ts = (time.perf_counter() - t) / len(seqs)
    assert strs == seqs
    t = time.perf_counter(); d = [a ^ a for a in objs]; tx = (time.perf_counter() - t) / len(seqs)
    assert not any(d)
    return {"pack_us": round(tp * 1e6, 2), "str_us": round(ts * 1e6, 2), "xor_us": round(tx * 1e6, 2)}


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
    bench(sq, 300)
    print("ours", bench(sq, n))
    try:
        from oracle import ref
        print("reference", bench(ref.load(), n))
    except Exception as e:  # the built reference is optional here
        print("reference unavailable:", e)
