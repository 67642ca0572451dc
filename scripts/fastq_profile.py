"""cProfile of shortseq_b200.read_and_count_fastq on a synthetic 2M-read FASTQ file (development aid)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import shortseq_b200 as sq
from fastq_bench import make_fastq

text = make_fastq(2_000_000, 2_000_000, 32)
path = "/tmp/ssq_prof.fastq"
text.tofile(path)
sq.read_and_count_fastq(path)
t0 = time.perf_counter()
g = sq.read_and_count_fastq(path)
print("second call", round(time.perf_counter() - t0, 3), "s", len(g))
cProfile.run("sq.read_and_count_fastq(path)", "/tmp/ssq_prof.out")
pstats.Stats("/tmp/ssq_prof.out").sort_stats("tottime").print_stats(14)
os.remove(path)
