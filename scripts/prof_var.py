"""Runs ShortSeqVar pack + decode on the C4 mix (150 / 300 / 1000 nt at 50 / 30 / 20 %) a few times and prints device
times (development aid; also the workload of the ncu captures of the Var kernels).  usage: prof_var.py [n_reads]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import shortseq_b200 as sq
from shortseq_b200 import _lib
from shortseq_b200._runtime import ptr


def timed(fn, iters=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(iters):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / iters


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    ctx = sq.pack_batch([b"ACGT" * 30], klass=sq.CLASS_VAR).ctx
    dev = ctx.device
    g = torch.Generator(device=dev)
    g.manual_seed(0x5EED0004)
    r = torch.rand(n, generator=g, device=dev)
    lens = torch.where(r < 0.5, 150, torch.where(r < 0.8, 300, 1000)).to(torch.int64)
    offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens, 0, out=offsets[1:])
    total = int(offsets[-1].item())
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    ascii_t = lut[torch.randint(0, 4, (total,), generator=g, device=dev, dtype=torch.uint8).long()]
    lib, h = _lib.lib(), ctx.bind()
    bound = lib.ssq_packvar_words_bound(total, n)
    words = ctx.empty((bound,), torch.int64)
    vlens = ctx.empty((n,), torch.int16)
    word_off = ctx.empty((n + 1,), torch.int64)
    out = ctx.empty((total,), torch.uint8)
    out_off = ctx.empty((n + 1,), torch.int64)

    def pack():
        _lib.check(lib.ssq_packvar(h, ptr(ascii_t), total, ptr(offsets), n, ptr(word_off), ptr(words), ptr(vlens)))

    def decode():
        _lib.check(lib.ssq_lens_to_offsets(h, ptr(vlens), 2, n, ptr(out_off)))
        _lib.check(lib.ssq_decodevar(h, ptr(words), ptr(word_off), ptr(vlens), n, ptr(out_off), ptr(out)))

    pm = timed(pack)
    dm = timed(decode)
    nwords = int(word_off[-1].item())
    pack_bytes = total + 8 * n + 8 * nwords + 8 * n + 2 * n
    decode_bytes = 8 * nwords + 2 * n + 8 * n + total + 8 * n
    ok = bool(torch.equal(out, ascii_t))
    print(f"var n={n} bases={total}: pack {pm:.3f} ms = {pack_bytes / pm / 1e6:.0f} GB/s, decode(+scan) {dm:.3f} ms = "
          f"{decode_bytes / dm / 1e6:.0f} GB/s, round trip ok={ok}, status={ctx.sync().code}")


if __name__ == "__main__":
    main()
