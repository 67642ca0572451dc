"""Wall time of the steps of read_and_count_fastq (development aid)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch
import shortseq_b200 as sq
from shortseq_b200 import _lib, counter as Cn
from shortseq_b200._runtime import context
from fastq_bench import make_fastq

text = make_fastq(2_000_000, 2_000_000, 32)
path = "/tmp/ssq_steps.fastq"
text.tofile(path)
lib = _lib.lib()
for rep_ in range(3):
    T = [time.perf_counter()]
    def tick(label):
        torch.cuda.synchronize()
        T.append(time.perf_counter()); print(f"  {label:28s} {1e3 * (T[-1] - T[-2]):8.1f} ms", flush=True)
    data = np.fromfile(path, dtype=np.uint8); tick("fromfile")
    ctx = context(None)
    ctr = sq.DeviceCounter(0, expected_unique=2_200_000); tick("create counter")
    nr, nl, fl, rep = C.c_int64(), C.c_int64(), C.c_int64(), _lib.Report()
    h = ctx.bind()
    _lib.check(lib.ssq_host_fastq_count(h, ctr.handle, None, data.ctypes.data, int(data.size), 0, 1, C.byref(nr), C.byref(nl), C.byref(fl), C.byref(rep))); tick("ssq_host_fastq_count")
    n = len(ctr); tick("len")
    keys, counts, first, _ = ctr.export(1, with_first_index=True); tick("export")
    w, l, _ = keys.to_host(); cnt = counts.cpu().numpy(); fi = first.cpu().numpy(); tick("to host")
    d = Cn.ShortSeqCounter(); Cn._fill_in_order(d, [(0, w, l, cnt, fi)]); tick("fill dict")
    hh = ctr.handle; ctr.handle = None; lib.ssq_counter_destroy(hh); tick("destroy counter")
    del keys, counts, first; tick("free tensors")
    del d; tick("free dict")
    print("total", round(1e3 * (T[-1] - T[0]), 1), "ms; status", rep.code, "reads", nr.value, "uniques", n)
os.remove(path)
