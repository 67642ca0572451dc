"""Multi-GPU dedup counting: local tables merged by a hash-partitioned all-to-all.

One process per GPU.  Each rank packs and counts its own shard of the reads with no communication; the only exchange
step moves UNIQUES, not reads, to their owner rank (owner = top log2(P) bits of the key hash) and adds them up there.
The global counter is the disjoint union of the P owner tables.  The reference has no multi-process mode (SURVEY
section 8e); this is new.

  Comm.merge        the product path: ssq_counter_merge_alltoall inside the C library (csrc/ssq_comm.cu)
  merge_alltoall    the same exchange spelled with torch.distributed collectives (export, all-to-all of the sizes,
                    all-to-all-v of words / lens / counts, weighted insert): works on any backend, which is what the
                    gloo tests on CPU exercise, and serves as bench.py --nccl-exchange
"""
import os
import time

import torch
import torch.distributed as dist

from ._lib import CLASS_64

_TIMING = os.environ.get("SSQ_MERGE_TIMING") == "1"      # development: print the wall time of each merge phase (adds syncs)


def _tick(label, t0, ctx):
    if not _TIMING:
        return t0
    torch.cuda.synchronize(ctx.device)
    t1 = time.perf_counter()
    if dist.get_rank() == 0:
        print(f"  [merge] {label}: {(t1 - t0) * 1e3:.2f} ms", flush=True)
    return t1


def exchange_counts(send_counts, group=None):
    """send_counts[p] = tuples this rank sends to rank p  ->  recv_counts[p] = tuples rank p sends here."""
    recv = torch.empty_like(send_counts)
    dist.all_to_all_single(recv, send_counts, group=group)
    return recv


def exchange_tuples(words, lens, counts, send_counts, recv_counts, group=None):
    """all-to-all-v of the exported tuples.  words [n] or [n, 3] int64, lens uint8, counts int64, all laid
    out as P consecutive segments of send_counts[p] tuples.  Works on any backend (NCCL on GPUs, gloo on CPU)."""
    send = [int(x) for x in send_counts.tolist()]
    recv = [int(x) for x in recv_counts.tolist()]
    total = sum(recv)
    out_words = words.new_empty((total,) + tuple(words.shape[1:]))
    out_lens = lens.new_empty((total,))
    out_counts = counts.new_empty((total,))
    dist.all_to_all_single(out_words, words.contiguous(), recv, send, group=group)
    dist.all_to_all_single(out_lens, lens.contiguous(), recv, send, group=group)
    dist.all_to_all_single(out_counts, counts.contiguous(), recv, send, group=group)
    return out_words, out_lens, out_counts


def merge_alltoall(local, group=None, owner=None):
    """Merge every rank's DeviceCounter `local` into per-rank owner tables.

    -> the DeviceCounter holding the keys this rank owns (global counts).  `owner` may be a
    pre-sized DeviceCounter created with hash_rot = log2(world) to reuse across calls.
    """
    from .counter import DeviceCounter
    world = dist.get_world_size(group)
    if world & (world - 1):
        raise ValueError("merge_alltoall needs a power-of-two world size")
    rot = world.bit_length() - 1
    t0 = _tick("start", time.perf_counter(), local.ctx) if _TIMING else 0.0
    keys, counts, _, parts = local.export(world)
    t0 = _tick("export", t0, local.ctx)
    if world == 1:
        recv_w, recv_l, recv_c = keys.words, keys.lens, counts
    else:
        recv_counts = exchange_counts(parts, group)
        t0 = _tick("exchange sizes", t0, local.ctx)
        recv_w, recv_l, recv_c = exchange_tuples(keys.words, keys.lens, counts, parts, recv_counts, group)
        t0 = _tick("exchange tuples", t0, local.ctx)
    if owner is None:
        owner = DeviceCounter(local.klass, expected_unique=int(recv_l.numel()), hash_rot=rot, device=local.ctx.device)
    owner.merge(recv_w, recv_l, recv_c)
    _tick("merge into owner table", t0, local.ctx)
    return owner


class Comm:
    """The library's own communicator (ssq_comm_*, csrc/ssq_comm.cu): NCCL loaded by the C library plus receive buffers
    shared between the single-GPU processes of the box through CUDA IPC.  `merge(local, owner)` is
    ssq_counter_merge_alltoall: sizes by ncclAllGather; ShortSeq64 counters are exchanged by the export kernel storing
    each owner's share straight into that owner's memory over NVLink, with device-side arrival flags instead of a host
    barrier; ShortSeq192 counters (and boxes without CUDA IPC) by grouped ncclSend / ncclRecv inside the library.
    torch.distributed only carries the 128-byte ncclUniqueId from rank 0 to the others -- a Cython / C host would pass
    it over whatever it has (MPI, a file)."""

    def __init__(self, ctx, group=None):
        import ctypes as C
        from . import _lib
        self.ctx, self.group = ctx, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        lib = _lib.lib()
        uid = (C.c_ubyte * 128)()
        if self.rank == 0:
            _lib.check(lib.ssq_comm_unique_id(uid))
        backend = dist.get_backend(group)
        t = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device=ctx.device if backend == "nccl" else "cpu")
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        uid = (C.c_ubyte * 128)(*t.cpu().tolist())
        h = C.c_void_p()
        _lib.check(lib.ssq_comm_init(ctx.bind(), uid, self.rank, self.world, C.byref(h)))
        self.handle = h
        self.exchange_ms, self.merge_ms = 0.0, 0.0
        self.streams, self._attached, self.last_streamed = False, None, False

    @property
    def peer_stores(self):
        """True when ShortSeq64 merges go over NVLink peer stores (CUDA IPC available on every rank)."""
        from . import _lib
        return bool(_lib.lib().ssq_comm_uses_peer_stores(self.handle))

    def attach(self, local, owner):
        """Collective: ssq_comm_attach -- from now on `local`'s counting passes stream every table region to its owner rank
        from inside the count kernel, and merge(local, owner) only publishes the arrival flags and merges.  Returns True
        when streaming is active (ShortSeq64 tables of one capacity on every rank, CUDA IPC available)."""
        import ctypes as C
        from . import _lib
        self.ctx.bind()
        on = C.c_int(0)
        _lib.check(_lib.lib().ssq_comm_attach(self.handle, local.handle if local is not None else None,
                                              owner.handle if owner is not None else None, C.byref(on)))
        self.streams = bool(on.value)
        self._attached = (local, owner)          # keep the counters alive while the library points at them
        return self.streams

    def merge(self, local, owner):
        """Collective: add every rank's `local` counter into the per-rank `owner` tables (owner must have been created
        with hash_rot = log2(world)).  Returns owner; self.exchange_ms / self.merge_ms hold this rank's device times."""
        import ctypes as C
        from . import _lib
        from . import batch as _batch
        self.ctx.bind()
        e, m = C.c_float(), C.c_float()
        _lib.check(_lib.lib().ssq_counter_merge_alltoall(self.handle, local.handle, owner.handle, C.byref(e), C.byref(m)))
        self.exchange_ms, self.merge_ms = e.value, m.value
        self.last_streamed = bool(_lib.lib().ssq_comm_last_merge_streamed(self.handle))
        _batch.raise_for_report(self.ctx.sync())
        return owner

    def close(self):
        """Collective."""
        from . import _lib
        h, self.handle = self.handle, None
        if h is not None and h.value:
            _lib.lib().ssq_comm_destroy(h)


def global_size(owner, group=None):
    """Number of distinct keys over all ranks (sum of the disjoint owner tables)."""
    t = torch.tensor([len(owner)], dtype=torch.int64, device=owner.ctx.device)
    dist.all_reduce(t, group=group)
    return int(t.item())
