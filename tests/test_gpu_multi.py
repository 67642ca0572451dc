"""GPU, multi-rank: the hash-partitioned all-to-all merge over NCCL (needs >= 2 GPUs; skipped otherwise)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, klass, out_dir, peer=False):
    import torch.distributed as dist
    import shortseq_b200 as sq
    from shortseq_b200.distributed import Comm, global_size, merge_alltoall
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        lo, hi = (18, 32) if klass == 0 else (40, 96)
        b = sq.synth_reads(300_000, 40_000, lo, hi, seed=0x5EED0001, first_read=rank * 300_000)
        local = sq.DeviceCounter(klass, expected_unique=50_000)
        local.pack_count(b)
        if peer:
            comm = Comm(local.ctx)
            owner = sq.DeviceCounter(klass, expected_unique=40_000, hash_rot=world.bit_length() - 1)
            for _ in range(3):                      # repeatedly: buffers and arrival flags are reused, the owner table is cleared in between
                owner.clear()
                comm.merge(local, owner)
            assert comm.peer_stores or os.environ.get("SSQ_NO_PEER_EXCHANGE")
            comm.close()
        else:
            owner = merge_alltoall(local)
        total = global_size(owner)
        keys, counts, _, _ = owner.export(1)
        w, l, _ = keys.to_host()
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), w=w, l=l, c=counts.cpu().numpy(), total=total)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("klass,peer", [(0, False), (1, False), (0, True), (1, True)])
def test_multi_gpu_merge_matches_oracle(tmp_path, klass, peer, oracle):
    """peer=True: ssq_counter_merge_alltoall inside the C library -- ShortSeq64: the export kernel stores each owner's tuples
    straight into that owner's memory (CUDA IPC + NVLink) and the owner waits for device-side arrival flags; ShortSeq192:
    the scatter pass of its two-pass export stores into the owners the same way (grouped ncclSend / ncclRecv only when
    CUDA IPC is unavailable).  peer=False: the torch.distributed spelling of the same exchange."""
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from shortseq_b200 import hashing
    mp.spawn(_worker, args=(world, _free_port(), klass, str(tmp_path), peer), nprocs=world, join=True)
    lo, hi = (18, 32) if klass == 0 else (40, 96)
    buf, off = oracle.synth_reads(0x5EED0001, 0, 600_000, 40_000, lo, hi)
    ow, ol, _ = oracle.pack_batch(klass, buf, off)
    uw, ul, uc, _ = oracle.count(ow, ol, 1 if klass == 0 else 3)
    expect = {(int(ul[j]), tuple(np.atleast_1d(uw[j]).tolist())): int(uc[j]) for j in range(len(ul))}
    merged = {}
    for r in range(world):
        d = np.load(tmp_path / f"r{r}.npz")
        assert int(d["total"]) == len(expect)
        assert (hashing.owner_rank(d["w"], d["l"], klass, world) == r).all()
        for j in range(len(d["l"])):
            key = (int(d["l"][j]), tuple(np.atleast_1d(d["w"][j]).tolist()))
            assert key not in merged
            merged[key] = int(d["c"][j])
    assert merged == expect


def _stream_worker(rank, world, port, out_dir):
    import faulthandler
    import torch.distributed as dist
    faulthandler.dump_traceback_later(int(os.environ.get("SSQ_TEST_HANG_S", "240")), exit=True)   # a hung collective fails the test instead of the box
    import shortseq_b200 as sq
    from shortseq_b200.distributed import Comm, global_size
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        n = 1_500_000
        b = sq.synth_reads(n, 1_200_000, 18, 32, seed=0x5EED0091, first_read=rank * n)
        local = sq.DeviceCounter(0, expected_unique=2_000_000)          # 2^22 slots: the region path
        owner = sq.DeviceCounter(0, expected_unique=1_600_000 // world, hash_rot=world.bit_length() - 1)
        comm = Comm(local.ctx)
        assert comm.attach(local, owner), "streaming should be active (CUDA IPC, nested region grids)"
        for rep in range(3):                    # both buffer parities, reused flags
            local.clear()
            owner.clear()
            local.pack_count(b)
            comm.merge(local, owner)
            assert comm.last_streamed
        keys, counts, _, _ = owner.export(1)
        w, l, _ = keys.to_host()
        np.savez(os.path.join(out_dir, f"a{rank}.npz"), w=w, l=l, c=counts.cpu().numpy(), total=global_size(owner))
        # a second pass into the populated table (regions loaded, earlier keys travel with their whole count): doubled counts
        local.pack_count(b)
        owner.clear()
        comm.merge(local, owner)
        assert comm.last_streamed
        keys, counts, _, _ = owner.export(1)
        w, l, _ = keys.to_host()
        np.savez(os.path.join(out_dir, f"b{rank}.npz"), w=w, l=l, c=counts.cpu().numpy(), total=global_size(owner))
        # the table changes after the streamed pass: the merge must not use the stale regions (unstreamed path, every rank alike)
        local.insert(sq.pack_batch(sq.synth_reads(1000, 1_200_000, 18, 32, seed=0x5EED0091, first_read=rank * n), klass=0))   # small: direct inserts
        owner.clear()
        comm.merge(local, owner)
        assert not comm.last_streamed
        keys, counts, _, _ = owner.export(1)
        w, l, _ = keys.to_host()
        np.savez(os.path.join(out_dir, f"c{rank}.npz"), w=w, l=l, c=counts.cpu().numpy(), total=global_size(owner))
        comm.attach(None, None)
        local.clear()
        owner.clear()
        local.pack_count(b)
        comm.merge(local, owner)
        assert not comm.last_streamed
        assert global_size(owner) == int(np.load(os.path.join(out_dir, f"a{rank}.npz"))["total"])
        comm.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2])
def test_streamed_exchange_matches_oracle(tmp_path, world, oracle):
    """ssq_comm_attach: the count kernel sends every table region to its owner rank while it counts; the merge only
    publishes the arrival flags and adds the fixed-place blocks up.  world = 1 runs the whole protocol on one GPU."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from shortseq_b200 import hashing
    mp.spawn(_stream_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    n = 1_500_000
    buf, off = oracle.synth_reads(0x5EED0091, 0, n * world, 1_200_000, 18, 32)
    ow, ol, _ = oracle.pack_batch(0, buf, off)
    uw, ul, uc, _ = oracle.count(ow, ol, 1)
    expect = dict(zip(zip(ul.tolist(), ((x,) for x in uw.tolist())), uc.tolist()))
    extra = {}
    for r in range(world):                      # the 1000 reads every rank inserted again before merge "c"
        lo_, hi_ = int(off[r * n]), int(off[r * n + 1000])
        xw, xl, _ = oracle.pack_batch(0, buf[lo_:hi_], off[r * n:r * n + 1001] - lo_)
        for j in range(1000):
            k = (int(xl[j]), (int(xw[j]),))
            extra[k] = extra.get(k, 0) + 1
    for tag, mult in (("a", 1), ("b", 2), ("c", 2)):
        merged = {}
        for r in range(world):
            d = np.load(tmp_path / f"{tag}{r}.npz")
            w, l, c = d["w"], d["l"], d["c"]                 # (an NpzFile re-reads the array at every access)
            assert int(d["total"]) == len(expect)
            assert (hashing.owner_rank(w, l, 0, world) == r).all()
            merged.update(zip(zip(l.tolist(), ((x,) for x in w.tolist())), c.tolist()))
        assert merged == {k: v * mult + (extra.get(k, 0) if tag == "c" else 0) for k, v in expect.items()}, tag
