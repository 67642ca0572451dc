"""CPU, world_size 2 (gloo): the host-side plumbing of the multi-GPU merge.

The GPU steps (export grouped by owner, weighted insert) are stood in for by the oracle; what is under
test is shortseq_b200.distributed's exchange (segment sizes, all-to-all-v of the tuples) and the owner
function.  The same functions run unchanged over NCCL on the GPU box (tests/test_gpu_multi.py, bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from shortseq_b200 import hashing
from shortseq_b200.distributed import exchange_counts, exchange_tuples


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shard(rank, klass):
    lo, hi = (15, 32) if klass == 0 else (33, 96)
    buf, off = O.synth_reads(0x5EED0001, rank * 20000, 20000, 3000, lo, hi)
    w, l, _ = O.pack_batch(klass, buf, off)
    return O.count(w, l, 1 if klass == 0 else 3)[:3], (w, l)


def _worker(rank, world, port, klass, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        (uw, ul, uc), _ = _shard(rank, klass)
        owner = hashing.owner_rank(uw, ul, klass, world)
        order = np.argsort(owner, kind="stable")                 # what the device export produces: grouped by owner
        send_counts = torch.from_numpy(np.bincount(owner, minlength=world).astype(np.int64))
        words = torch.from_numpy(np.ascontiguousarray(uw[order]).view(np.int64))
        lens = torch.from_numpy(ul[order].astype(np.uint8))
        counts = torch.from_numpy(uc[order].astype(np.int64))
        recv_counts = exchange_counts(send_counts)
        rw, rl, rc = exchange_tuples(words, lens, counts, send_counts, recv_counts)
        assert rl.numel() == int(recv_counts.sum())
        rw_np = rw.numpy().view(np.uint64)
        assert (hashing.owner_rank(rw_np, rl.numpy(), klass, world) == rank).all(), "received a key this rank does not own"
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), w=rw_np, l=rl.numpy(), c=rc.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("klass", [0, 1])
def test_alltoall_merge_two_ranks(tmp_path, klass):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), klass, str(tmp_path)), nprocs=world, join=True)
    merged = {}
    for r in range(world):
        d = np.load(tmp_path / f"r{r}.npz")
        for j in range(len(d["l"])):
            w = (int(d["w"][j]),) if klass == 0 else tuple(int(x) for x in d["w"][j])
            key = (int(d["l"][j]), w)
            merged[key] = merged.get(key, 0) + int(d["c"][j])
    # expectation: one oracle count over both shards
    ws, ls = zip(*[_shard(r, klass)[1] for r in range(world)])
    uw, ul, uc, _ = O.count(np.concatenate(ws), np.concatenate(ls), 1 if klass == 0 else 3)
    expect = {(int(ul[j]), (int(uw[j]),) if klass == 0 else tuple(int(x) for x in uw[j])): int(uc[j]) for j in range(len(ul))}
    assert merged == expect
