#!/usr/bin/env python3
"""Generate tests/golden/ref_vectors.json from the UNMODIFIED reference.

Run in the build container (needs oracle/_ref, built by oracle/build_ref.py from
/root/reference):   python tests/golden/make_golden.py

Every value is produced by the reference itself: packed words are read out of the
reference objects' memory (oracle/ref.py), hashes by hash(), strings by str(),
Hamming distances by `^`, counts by ShortSeqCounter.  The GPU box has no
/root/reference; tests there use this file.
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref as R  # noqa: E402

KATS = ["A", "C", "T", "G", "ACGT", "ATGC", "GATTACA", "TGACTGACTGAC", "TGAGGTAGTAGGTTGTATAGTT",
        "GCGTAATAGGGGGTTTCGCTGTGGGGCGGCTAG", "G" * 32, "ACGT" * 8, "A" * 32 + "C",
        "TATTAGCGATTGACAGTTGTCCTGTAATAACGCCGGGTAAATTTGCCG", "TATTACCGATTGACAGTTGTCCTGTAATAACGGCGGGTAAATTTGCTG",
        "ACGT" * 24, "ACGT" * 24 + "T", "", "A" * 1024, "T" * 96, "C" * 33]


def rnd(rng, n):
    return "".join(rng.choice("ACGT") for _ in range(n))


def main():
    sq = R.load()
    assert sq is not None, "build oracle/_ref first (python oracle/build_ref.py)"
    rng = random.Random(20261018)
    seqs = list(KATS)
    for L in list(range(0, 131)) + [150, 255, 256, 257, 300, 511, 512, 513, 1000, 1023, 1024]:
        seqs.append(rnd(rng, L))
    pack = []
    for s in seqs:
        o = sq.pack(s)
        pack.append({"seq": s, "type": type(o).__name__, "len": len(o),
                     "words": [f"{w:#x}" for w in R.raw_words(o)], "hash": hash(o), "str": str(o)})
    ham = []
    for L in [1, 2, 12, 22, 31, 32, 33, 50, 64, 65, 95, 96, 97, 128, 129, 300, 1024]:
        for _ in range(3):
            a = rnd(rng, L)
            b = list(a)
            for _ in range(rng.randrange(0, min(L, 9) + 1)):
                b[rng.randrange(L)] = rng.choice("ACGT")
            b = "".join(b)
            ham.append({"a": a, "b": b, "dist": sq.pack(a) ^ sq.pack(b)})
    counters = []
    cases = [[b"ACGT", b"TTTT", b"ACGT", b"GG", b"TTTT", b"ACGT"], [b"A", b"AA", b"AAA", b"A"], [b"", b"", b"A"],
             [b"ATGC"] * 10]
    pool = [rnd(rng, rng.randrange(15, 33)).encode() for _ in range(40)]
    cases.append([rng.choice(pool) for _ in range(400)])
    pool = [rnd(rng, rng.randrange(33, 97)).encode() for _ in range(30)]
    cases.append([rng.choice(pool) for _ in range(300)])
    pool = [rnd(rng, rng.randrange(10, 80)).encode() for _ in range(30)]   # mixed 64 / 192
    cases.append([rng.choice(pool) for _ in range(300)])
    for reads in cases:
        c = sq.ShortSeqCounter(reads)
        counters.append({"reads": [r.decode() for r in reads],
                         "items": [[str(k), type(k).__name__, v] for k, v in c.items()]})
    rejects = []
    for s in ["N", "*", "U", "a", "acgt", "A" * 32 + "N", "A" * 8 + "N" + "A" * 25, "ACGN", "A" * 100 + "N" + "A" * 30,
              "A" * 1025]:
        try:
            sq.pack(s)
            raise AssertionError(f"reference accepted {s!r}")
        except Exception as e:  # noqa: BLE001
            rejects.append({"seq": s, "message": str(e)})
    slices = []
    for L in [10, 32, 33, 64, 96, 97, 200, 1024]:
        s = rnd(rng, L)
        o = sq.pack(s)
        for _ in range(12):
            a = rng.randrange(0, L)
            b = rng.randrange(a, L + 1)
            sl = o[a:b]
            slices.append({"seq": s, "start": a, "stop": b, "type": type(sl).__name__,
                           "words": [f"{w:#x}" for w in R.raw_words(sl)] if b > a else ["0x0"], "hash": hash(sl)})
    out = {"generator": "tests/golden/make_golden.py", "reference": "AlexTate/ShortSeq (unmodified, built by oracle/build_ref.py)",
           "pack": pack, "hamming": ham, "counters": counters, "rejects": rejects, "slices": slices}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print(path, {k: len(v) for k, v in out.items() if isinstance(v, list)})


if __name__ == "__main__":
    main()
